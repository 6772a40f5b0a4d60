"""Statistical equivalence of the "--rng=philox" path with the reference's own release binary over 20 seeds
(north_star: KS and chi-square at p > 0.01 for the stochastic outputs).

The CUDA path reproduces the oracle's philox mode bit for bit (tests/test_gpu_parity.py), so the comparison
binary <-> philox spec can run on the CPU: 20 seeds of the release binary against 20 seeds of the spec.  The
`gpu`-marked variant repeats it with the CUDA path itself on the GPU box.  Skipped where the binary is absent.

No retries: the family of statistics is judged once, on N_SEEDS runs per arm.  With m statistics at the 0.01 level one
of them falls below 0.01 by chance in about one run out of five (the binary's PCR and sampling stages are not reproducible
run to run - Go map order, SURVEY 7.2 - so the p-values are random even for fixed seeds), hence the family-wise form:
every p-value must exceed 0.01 / m (Bonferroni, family-wise level 1 %) and at most two may fall below 0.01.

test_pcr_copy_number_law_per_stratum pins the law of AmplifyFragment (thermocycler.go:133-147, random.go:168-233) per
(efficiency, cycles) stratum against the binary: with a fixed efficiency -e every fragment is an independent
Galton-Watson lineage, so the per-length amplification ratio has a known mean and variance.
"""
import numpy as np
import pytest
from scipy import stats

import util

ARGV = ["-n", "8000", "-m", "0.25", "-c", "11"]  # C1 at a quarter of the expression: 5890 molecules per run
N_SEEDS = 24


def _refbin():
    from oracle import refbin
    if not refbin.available():
        import subprocess, os
        subprocess.call(["make", "-s", "-C", os.path.join(os.path.dirname(util.HERE), "oracle"), "ref"])
    if not refbin.available():
        pytest.skip("reference release binary not available (oracle/_ref)")
    return refbin


def _gc_percent(seq):
    a = np.frombuffer(seq, dtype=np.uint8)
    return int(100.0 * np.count_nonzero((a == ord("G")) | (a == ord("C"))) / max(len(a), 1))


def _collect_binary(refbin, seeds):
    acc = dict(frag={}, pcr_ratio=[], per_tr={}, gc={}, sampled_len={}, minus=0, n=0, runs=[])
    for s in seeds:
        acc["runs"].append(dict(gc=[], tr={}))
        r = refbin.run(ARGV + ["-si", str(s)], util.BASIC_FASTA, keep_fasta=True)
        rep = r["report"]
        for k, v in rep["Fragdist after fragmentation"]["Data"].items():
            acc["frag"][int(k)] = acc["frag"].get(int(k), 0) + v
        nf = sum(rep["Fragdist after fragmentation"]["Data"].values())
        acc["pcr_ratio"].append(sum(rep["Fragdist after PCR"]["Data"].values()) / nf)
        for fid, name, strand, start, end, seq in refbin.parse_fasta(r["fasta"]):
            acc["per_tr"][name] = acc["per_tr"].get(name, 0) + 1
            g = _gc_percent(seq)
            acc["gc"][g] = acc["gc"].get(g, 0) + 1
            acc["runs"][-1]["gc"].append(g)
            acc["runs"][-1]["tr"][name] = acc["runs"][-1]["tr"].get(name, 0) + 1
            acc["sampled_len"][end - start] = acc["sampled_len"].get(end - start, 0) + 1
            acc["minus"] += strand == "-"
            acc["n"] += 1
    return acc


def _collect_ours(run_one, seeds):
    acc = dict(frag={}, pcr_ratio=[], per_tr={}, gc={}, sampled_len={}, minus=0, n=0, runs=[])
    for s in seeds:
        acc["runs"].append(dict(gc=[], tr={}))
        args = util.parse(ARGV + ["-si", str(s)])
        names, seqs, levels = util.transcripts(args)
        frag, pool_total, out, fasta = run_one(args, names, seqs, levels)
        for k, v in frag.items():
            acc["frag"][k] = acc["frag"].get(k, 0) + v
        acc["pcr_ratio"].append(pool_total / sum(frag.values()))
        lines = fasta.split(b"\n")
        for i, t in enumerate(out["tr"]):
            nm = names[int(t)]
            acc["per_tr"][nm] = acc["per_tr"].get(nm, 0) + 1
            g = _gc_percent(lines[2 * i + 1])
            acc["gc"][g] = acc["gc"].get(g, 0) + 1
            acc["runs"][-1]["gc"].append(g)
            acc["runs"][-1]["tr"][nm] = acc["runs"][-1]["tr"].get(nm, 0) + 1
            l = int(out["end"][i] - out["start"][i])
            acc["sampled_len"][l] = acc["sampled_len"].get(l, 0) + 1
        acc["minus"] += int(out["minus"].sum())
        acc["n"] += len(out["tr"])
    return acc


def _chi2(a, b, min_count=10):
    keys = sorted(set(a) | set(b))
    xs, ys, cx, cy = [], [], 0.0, 0.0
    for k in keys:
        cx += a.get(k, 0); cy += b.get(k, 0)
        if cx + cy >= 2 * min_count:
            xs.append(cx); ys.append(cy); cx = cy = 0.0
    if cx + cy > 0 and xs:
        xs[-1] += cx; ys[-1] += cy
    return stats.chi2_contingency(np.array([xs, ys]))[1]


def _expand(h):
    return np.repeat(np.array(sorted(h)), [h[k] for k in sorted(h)])


def _compare(ref, ours):
    """Molecule-level outputs (fragmentation) are iid across molecules -> pooled chi-square / KS.  Outputs downstream
    of PCR are over-dispersed (a few jackpot species dominate a pool; the reference compared with itself on other
    seeds fails a pooled chi-square at p ~ 1e-20), so they are compared as per-run statistics, 20 runs against 20."""
    p = {}
    p["fragment length after fragmentation (chi2)"] = _chi2(ref["frag"], ours["frag"])
    p["fragment length after fragmentation (KS)"] = stats.ks_2samp(_expand(ref["frag"]), _expand(ours["frag"]))[1]
    p["sampled fragment length = target draw (chi2)"] = _chi2(ref["sampled_len"], ours["sampled_len"])
    p["PCR copy number: amplification factor per run (KS)"] = stats.ks_2samp(ref["pcr_ratio"], ours["pcr_ratio"])[1]
    p["PCR copy number: mean (Welch t)"] = stats.ttest_ind(ref["pcr_ratio"], ours["pcr_ratio"], equal_var=False)[1]
    p["PCR copy number: variance (Levene)"] = stats.levene(ref["pcr_ratio"], ours["pcr_ratio"])[1]
    gm = lambda a: [float(np.mean(r["gc"])) for r in a["runs"]]
    p["GC profile: mean GC% of sampled fragments per run (KS)"] = stats.ks_2samp(gm(ref), gm(ours))[1]
    p["GC profile: mean GC% per run (Welch t)"] = stats.ttest_ind(gm(ref), gm(ours), equal_var=False)[1]
    for lo, hi in ((0, 36), (36, 44), (44, 52), (52, 101)):
        fr = lambda a: [float(np.mean((np.array(r["gc"]) >= lo) & (np.array(r["gc"]) < hi))) for r in a["runs"]]
        p["GC profile: fraction in [%d,%d) per run (KS)" % (lo, hi)] = stats.ks_2samp(fr(ref), fr(ours))[1]
    for name in sorted(ref["per_tr"]):
        fr = lambda a: [r["tr"].get(name, 0) / max(sum(r["tr"].values()), 1) for r in a["runs"]]
        p["per-transcript sampled fraction %s per run (KS)" % name.decode()] = stats.ks_2samp(fr(ref), fr(ours))[1]
    tab = np.array([[ref["minus"], ref["n"] - ref["minus"]], [ours["minus"], ours["n"] - ours["minus"]]])
    p["strand (chi2)"] = stats.chi2_contingency(tab)[1]
    return p


def _equivalence(run_one):
    refbin = _refbin()
    p = _compare(_collect_binary(refbin, range(1, 1 + N_SEEDS)), _collect_ours(run_one, range(1001, 1001 + N_SEEDS)))
    for k, v in sorted(p.items(), key=lambda kv: kv[1]):
        print("   p = %.4f  %s" % (v, k))
    m = len(p)
    assert all(v > 0.01 / m for v in p.values()), {k: v for k, v in p.items() if not v > 0.01 / m}   # family-wise 1 %
    assert sum(1 for v in p.values() if not v > 0.01) <= 2, {k: v for k, v in p.items() if not v > 0.01}


STRATA = [(0.3, 5), (0.3, 11), (0.6, 8), (0.9, 5), (0.9, 11), (0.5, 14)]  # (fixed efficiency -e, cycles -c)


def _gw_moments(eff, cycles):
    """Mean and variance of the copy number of ONE fragment after `cycles` rounds of c <- c + Binomial(c, eff):
    a Galton-Watson process with offspring 1 + Bernoulli(eff)."""
    m, s2 = 1.0 + eff, eff * (1.0 - eff)
    return m ** cycles, s2 * m ** (cycles - 1) * (m ** cycles - 1.0) / (m - 1.0)


def _ratio_z(after_frag, after_pcr, eff, cycles, min_frags=30):
    """Standardised amplification ratio of every length class with enough fragments: z ~ N(0, 1) under the law."""
    mean, var = _gw_moments(eff, cycles)
    z = []
    for l, n in after_frag.items():
        if n >= min_frags:
            z.append((after_pcr.get(l, 0) / n - mean) / np.sqrt(var / n))
    return z


def test_pcr_copy_number_law_per_stratum(oracle):
    """PCR copy-number mean and variance per (efficiency, cycles) stratum.  With a fixed efficiency -e every fragment is an
    independent branching lineage, so the amplification ratio of a length class with n fragments is
    (1+e)^c +- sqrt(Var / n) with a known Var: a far sharper instrument than 20 per-run ratios.

    * the philox spec (BINV / BTRS, = the CUDA path bit for bit) must follow the exact law: mean (t) and variance (chi2) of
      the standardised ratios per stratum, family-wise level 1 %;
    * the release binary is held to an effect size, not a p-value: its sampler (Bernoulli sums below 25 copies, the NR
      rejection method run with p = max(pp, 1-pp) above: random.go:168-233) deviates from the exact law by up to ~0.8 % in the
      mean amplification (measured: -0.8 % at e=0.3 c=5, +0.5 % at e=0.3 c=11, -0.8 % at e=0.1 c=11; 4-6 sigma with ~1 750
      length classes) - invisible to the 20-seed KS / chi-square tests north_star prescribes, visible here.  The spec does
      not reproduce that deviation; both stay within 1.5 % of the exact mean and 15 % of the exact standard deviation."""
    refbin = _refbin()
    pvals, effect = {}, {}
    for eff, cycles in STRATA:
        zr, zo, fr, fo = [], [], [0, 0], [0, 0]
        for seed in range(1, 5):
            argv = ["-n", "100", "-e", str(eff), "-c", str(cycles)]
            r = refbin.run(argv + ["-si", str(seed)], util.BASIC_FASTA, keep_fasta=False)["report"]
            af, ap = util.panel_int(r["Fragdist after fragmentation"]["Data"]), util.panel_int(r["Fragdist after PCR"]["Data"])
            zr += _ratio_z(af, ap, eff, cycles)
            fr[0] += sum(af.values()); fr[1] += sum(ap.values())
            args = util.parse(argv + ["-si", str(100 + seed)])
            names, seqs, levels = util.transcripts(args)
            o = util.run_oracle(oracle, args, names, seqs, levels, oracle.MODE_PHILOX, stages="pcr")
            af, ap = o.hist_dict(oracle.H_AFTER_FRAG), o.hist_dict(oracle.H_AFTER_PCR)
            zo += _ratio_z(af, ap, eff, cycles)
            fo[0] += sum(af.values()); fo[1] += sum(ap.values())
        assert len(zr) > 200 and len(zo) > 200
        tag = "e=%g c=%d" % (eff, cycles)
        mean = _gw_moments(eff, cycles)[0]
        pvals[tag + " spec mean (t)"] = stats.ttest_1samp(zo, 0.0)[1]
        q = np.sum(np.square(zo))
        pvals[tag + " spec variance (chi2)"] = 2 * min(stats.chi2.cdf(q, len(zo)), stats.chi2.sf(q, len(zo)))
        effect[tag] = dict(spec_mean_rel=fo[1] / fo[0] / mean - 1.0, binary_mean_rel=fr[1] / fr[0] / mean - 1.0,
                           spec_z_sd=float(np.std(zo)), binary_z_sd=float(np.std(zr)), binary_z_mean=float(np.mean(zr)))
    for k, v in sorted(pvals.items(), key=lambda kv: kv[1]):
        print("   p = %.4f  %s" % (v, k))
    for k, v in effect.items():
        print("   %s: %s" % (k, {a: round(b, 4) for a, b in v.items()}))
    m = len(pvals)
    assert all(v > 0.01 / m for v in pvals.values()), {k: v for k, v in pvals.items() if not v > 0.01 / m}
    assert sum(1 for v in pvals.values() if not v > 0.01) <= 2, {k: v for k, v in pvals.items() if not v > 0.01}
    for k, v in effect.items():
        assert abs(v["spec_mean_rel"]) < 0.015 and abs(v["binary_mean_rel"]) < 0.015, (k, v)
        assert 0.85 < v["spec_z_sd"] < 1.15 and 0.85 < v["binary_z_sd"] < 1.15, (k, v)


def test_philox_spec_vs_release_binary(oracle):
    def run_one(args, names, seqs, levels):
        o = util.run_oracle(oracle, args, names, seqs, levels, oracle.MODE_PHILOX)
        return o.hist_dict(oracle.H_AFTER_FRAG), o.pool_total, o.sampled(), o.fasta(names)
    _equivalence(run_one)


@pytest.mark.gpu
def test_cuda_path_vs_release_binary(gpu):
    def run_one(args, names, seqs, levels):
        h = util.run_gpu(gpu, args, names, seqs, levels)
        return h.hist_dict(gpu.H_AFTER_FRAG, 400), int(h.stats().pool_total), h.sampled(), h.fasta(names)
    _equivalence(run_one)
