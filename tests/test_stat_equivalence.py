"""Statistical equivalence of the "--rng=philox" path with the reference's own release binary over 20 seeds
(north_star: KS and chi-square at p > 0.01 for the stochastic outputs).

The CUDA path reproduces the oracle's philox mode bit for bit (tests/test_gpu_parity.py), so the comparison
binary <-> philox spec can run on the CPU: 20 seeds of the release binary against 20 seeds of the spec.  The
`gpu`-marked variant repeats it with the CUDA path itself on the GPU box.  Skipped where the binary is absent.

Because the binary's PCR and sampling stages are not reproducible run to run (Go map order), a test whose
p-value falls below 0.01 is repeated once on 20 fresh seeds and must then pass (false-alarm rate 1e-4 instead of 1e-2).
"""
import numpy as np
import pytest
from scipy import stats

import util

ARGV = ["-n", "8000", "-m", "0.25", "-c", "11"]  # C1 at a quarter of the expression: 5890 molecules per run
N_SEEDS = 20


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
    failing = None
    for attempt in range(2):
        seeds_ref = range(1 + 100 * attempt, 1 + 100 * attempt + N_SEEDS)
        seeds_ours = range(1001 + 100 * attempt, 1001 + 100 * attempt + N_SEEDS)
        p = _compare(_collect_binary(refbin, seeds_ref), _collect_ours(run_one, seeds_ours))
        print("attempt %d:" % attempt)
        for k, v in p.items():
            print("   p = %.4f  %s" % (v, k))
        low = {k for k, v in p.items() if not v > 0.01}
        failing = low if failing is None else failing & low  # a statistic must fail on both seed sets to count
        if not failing:
            return
    assert not failing, failing


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
