"""The "--rng=philox" specification as restated by the oracle: building blocks and invariants (CPU only)."""
import ctypes
import math

import numpy as np
import pytest
from scipy import special, stats

import util


def test_philox_known_answers(oracle):
    L = oracle.lib()

    def px(c, k):
        c, k, o = np.array(c, np.uint32), np.array(k, np.uint32), np.zeros(4, np.uint32)
        L.orc_philox(c.ctypes.data, k.ctypes.data, o.ctypes.data)
        return list(o)

    # Random123 kat_vectors, philox4x32-10
    assert px([0] * 4, [0] * 2) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert px([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert px([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_go_rand_replica_first_values(oracle):
    """Go's documented stream: rand.New(rand.NewSource(1)).Float64() begins 0.6046602879796196, 0.9405090880450124,
    0.6645600532184904 (the values every Go programmer sees from the unseeded global generator)."""
    out = np.zeros(3)
    oracle.lib().orc_go_rand_floats(1, 3, out.ctypes.data)
    assert list(out) == [0.6046602879796196, 0.9405090880450124, 0.6645600532184904]


def test_detmath_accuracy(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(1)
    for x in np.concatenate([rng.uniform(-700, 700, 3000), rng.uniform(-1, 1, 3000)]):
        assert math.isclose(L.orc_det_exp(float(x)), math.exp(x), rel_tol=4.5e-16)
    for x in np.exp(rng.uniform(-700, 700, 6000)):
        assert math.isclose(L.orc_det_log(float(x)), math.log(x), rel_tol=4.5e-16, abs_tol=1e-300)
    ps = np.concatenate([rng.uniform(0, 1, 4000), 10.0 ** rng.uniform(-300, -1, 500)])
    for p in ps:
        assert math.isclose(L.orc_det_ndtri(float(p)), special.ndtri(p), rel_tol=2e-15, abs_tol=1e-15)
    for k in [0, 1, 2, 9, 10, 11, 100, 12345, 1e9, 1e15]:
        assert math.isclose(L.orc_det_logfact(float(k)), special.gammaln(k + 1), rel_tol=1e-11, abs_tol=1e-12)
    assert L.orc_det_exp(-1000.0) == 0.0 and L.orc_det_exp(1000.0) == math.inf and L.orc_det_log(0.0) == -math.inf


APPENDIX_C = {  # derived from nnthermo.go:68-170 at T = 5, k = 6
    b"AAAAAA": 27.420537338477043, b"TTTTTT": 27.420537338477043, b"GGGGGG": 52.03896441320523,
    b"ACGTAC": 56.17841501711723, b"CGCGCG": 161.4976741063506, b"GAATTC": 45.96815493765271,
    b"ACCCCA": 36.74549747563298, b"GCATGC": 75.75817022252053, b"NNNNNN": 0.9765769881895615,
    b"ACGNAC": 11.94693834674044,
}


def test_kmer_affinity_known_answers(oracle):
    """Affinities within 1e-6 relative (north_star); in fact ~1e-15 with either exp."""
    L = oracle.lib()
    for kmer, want in APPENDIX_C.items():
        for det in (0, 1):
            assert math.isclose(L.orc_kmer_affinity(kmer, 6, 5.0, det), want, rel_tol=1e-12)
    tot = 0.0
    for c in range(4096):
        kmer = bytes(b"ACGT"[(c >> (2 * i)) & 3] for i in range(6))
        tot += L.orc_kmer_affinity(kmer, 6, 5.0, 1)
    assert math.isclose(tot, 201711.76817515807, rel_tol=1e-12)


def test_binding_profile_quirks(oracle):
    """len-k entries (not len-k+1); reverse profile = profile of RevComp; non-ACGT -> 'N' on the minus strand."""
    L = oracle.lib()
    seq = b"ACGTRACGTTTGCANNACGT"
    n = len(seq) - 6
    f, r = np.zeros(n), np.zeros(n)
    L.orc_binding_profile(seq, len(seq), 6, 5.0, 0, 0, f.ctypes.data)
    L.orc_binding_profile(seq, len(seq), 6, 5.0, 1, 0, r.ctypes.data)
    rc = np.zeros(len(seq), dtype=np.uint8)
    L.orc_revcomp(seq, len(seq), rc.ctypes.data)
    rcs = rc.tobytes()
    assert rcs == b"ACGTNNTGCAAACGTNACGT"
    for i in range(n):
        assert f[i] == L.orc_kmer_affinity(seq[i:i + 6], 6, 5.0, 0)
        assert r[i] == L.orc_kmer_affinity(rcs[i:i + 6], 6, 5.0, 0)


def test_quantized_weights(oracle):
    L = oracle.lib()
    seq = b"CGCGCGCGAAAAAATTTNACGTACGTAC"
    n = len(seq) - 6
    w = np.zeros(n, dtype=np.uint64)
    L.orc_quantized_weights(seq, len(seq), 6, 5.0, 0, w.ctypes.data)
    assert w[0] == 67108863  # CGCGCG is the strongest hexamer: saturates at 2^26 - 1
    k = np.array([L.orc_kmer_affinity(seq[i:i + 6], 6, 5.0, 1) for i in range(n)])
    assert np.all(np.abs(w / 67108864.0 - k / k.max()) <= 2.0 ** -26)


@pytest.mark.parametrize("n,p", [(1, 0.5), (7, 0.3), (24, 0.9), (25, 0.3), (1000, 0.009), (1000, 0.0101), (1000, 0.75),
                                 (10**6, 0.5), (10**9, 0.37), (3 * 10**11, 0.8)])
def test_px_binomial_distribution(oracle, n, p):
    """One amplification cycle c + Binomial(c, p): chi-square (small n) or z-tests on mean and variance."""
    L = oracle.lib()
    m = 20000
    x = np.array([L.orc_px_binomial(99, i, n & 0xffffffff, 5, n, p, 1) - n for i in range(m)], dtype=np.float64)
    assert x.min() >= 0 and x.max() <= n
    if n <= 1000:
        ks = np.arange(0, n + 1)
        pmf = stats.binom.pmf(ks, n, p)
        keep = pmf * m >= 5
        obs = np.array([np.sum(x == k) for k in ks[keep]], dtype=np.float64)
        exp = pmf[keep] * m
        obs = np.append(obs, m - obs.sum())
        exp = np.append(exp, m - exp.sum())
        if exp[-1] < 1e-9:
            obs, exp = obs[:-1], exp[:-1]
        chi2 = ((obs - exp) ** 2 / np.maximum(exp, 1e-12)).sum()
        assert stats.chi2.sf(chi2, len(obs) - 1) > 1e-4
    mu, var = n * p, n * p * (1 - p)
    assert abs(x.mean() - mu) < 5 * math.sqrt(var / m)
    assert abs(x.var() - var) < 6 * var * math.sqrt(2.0 / m) + 1e-9


def test_px_binomial_edges(oracle):
    L = oracle.lib()
    assert L.orc_px_binomial(1, 0, 0, 0, 1000, 0.0, 11) == 1000          # p = 0: no growth
    assert L.orc_px_binomial(1, 0, 0, 0, 3, 1.0, 11) == 3 * 2048         # p = 1: exact doubling
    assert L.orc_px_binomial(1, 0, 0, 0, 0, 0.7, 5) == 0


@pytest.mark.parametrize("mean", [0.3, 4.0, 11.99, 12.0, 35.0, 600.0])
def test_px_poisson_distribution(oracle, mean):
    L = oracle.lib()
    m = 20000
    x = np.array([L.orc_px_poisson(3, 1, i, mean) for i in range(m)], dtype=np.float64)
    hi = int(stats.poisson.ppf(1 - 1e-7, mean)) + 2
    ks = np.arange(0, hi)
    pmf = stats.poisson.pmf(ks, mean)
    keep = pmf * m >= 5
    obs = np.array([np.sum(x == k) for k in ks[keep]], dtype=np.float64)
    exp = pmf[keep] * m
    obs, exp = np.append(obs, m - obs.sum()), np.append(exp, m - exp.sum())
    chi2 = ((obs - exp) ** 2 / np.maximum(exp, 1e-9)).sum()
    assert stats.chi2.sf(chi2, len(obs) - 1) > 1e-4


def _two_sample_chi2(a, b, min_count=10):
    keys = sorted(set(a) | set(b))
    x = np.array([a.get(k, 0) for k in keys], dtype=np.float64)
    y = np.array([b.get(k, 0) for k in keys], dtype=np.float64)
    # merge sparse neighbouring bins
    xs, ys, cx, cy = [], [], 0.0, 0.0
    for u, v in zip(x, y):
        cx += u; cy += v
        if cx + cy >= 2 * min_count:
            xs.append(cx); ys.append(cy); cx = cy = 0.0
    if cx + cy > 0 and xs:
        xs[-1] += cx; ys[-1] += cy
    return stats.chi2_contingency(np.array([xs, ys]))[1]


@pytest.mark.parametrize("d", ["1.0:sn:(189, 24, -1.09975, 76, 294)", "0.9:sn:(600,50,4,300,2000) + 0.1:n:(700,1,600,2000)",
                               "1:n:(300,20,100,5000)", "0.5:g:(150,1,0,220) + 0.5:g:(80,4,10,400)", "1:g:(60,0.5,5,300)"])
def test_target_lengths_philox_vs_go_stream(oracle, d):
    """a19: the spec's variates (inverse-CDF normal, Marsaglia-Tsang gamma) against the reference's (ziggurat,
    Knuth / Kundu-Gupta), 200k draws each."""
    args = util.parse(["-n", "200000", "-d", d, "-si", "5"])
    names, seqs, levels = util.transcripts(args)
    a = util.run_oracle(oracle, args, names, seqs, levels, oracle.MODE_REF, stages="target").hist_dict(oracle.H_TARGET)
    b = util.run_oracle(oracle, args, names, seqs, levels, oracle.MODE_PHILOX, stages="target").hist_dict(oracle.H_TARGET)
    assert sum(a.values()) == sum(b.values()) == 200000
    assert _two_sample_chi2(a, b) > 0.01


@pytest.mark.parametrize("name", ["c1", "after_prim", "after_noprim_double", "pre_prim", "prim_jump_300", "loss_k8"])
def test_fragmentation_philox_vs_go_stream(oracle, name):
    """Same model, two random streams: fragment-length, poly-A and per-transcript distributions must agree
    (chi-square, p > 0.01); this is the CPU side of the statistical-equivalence requirement."""
    args = util.parse(util.PIN_CONFIGS[name] + ["-si", "21"])
    names, seqs, levels = util.transcripts(args)
    a = util.run_oracle(oracle, args, names, seqs, levels, oracle.MODE_REF, stages="fragment")
    b = util.run_oracle(oracle, args, names, seqs, levels, oracle.MODE_PHILOX, stages="fragment")
    assert _two_sample_chi2(a.hist_dict(oracle.H_AFTER_FRAG), b.hist_dict(oracle.H_AFTER_FRAG)) > 0.01
    assert _two_sample_chi2(a.hist_dict(oracle.H_POLYA), b.hist_dict(oracle.H_POLYA)) > 0.01
    sa, sb = a.species(), b.species()
    per_tr = lambda s: {int(t): int(c) for t, c in zip(*np.unique(np.repeat(s["tr"], s["count0"].astype(np.int64)), return_counts=True))}
    assert _two_sample_chi2(per_tr(sa), per_tr(sb), 5) > 0.01
    # start-position profile along the most expressed transcript (priming bias), 40 bins
    pos = lambda s: {int(k): int(v) for k, v in zip(*np.unique(s["start"][s["tr"] == 0] // 30, return_counts=True))}
    assert _two_sample_chi2(pos(sa), pos(sb)) > 0.01


def test_selection_invariants(oracle):
    """sampler.go:65-106 semantics of the spec's draw: never more than the pool holds, missing = req - drawn,
    drained classes are taken whole, and the draw is without replacement."""
    args = util.parse(["-n", "60000", "-c", "2", "-d", "1:n:(150,3,140,160)", "-si", "5"])
    names, seqs, levels = util.transcripts(args)
    o = util.run_oracle(oracle, args, names, seqs, levels, oracle.MODE_PHILOX, stages="pcr")
    before = o.species()
    pcr = o.hist_dict(oracle.H_AFTER_PCR)
    o.sample()
    tgt, smp, mis, aft = (o.hist_dict(k) for k in (oracle.H_TARGET, oracle.H_SAMPLED, oracle.H_MISSING, oracle.H_AFTER_SAMPLING))
    for l, req in tgt.items():
        t = pcr.get(l, 0)
        assert smp.get(l, 0) == min(req, t) and mis.get(l, 0) == req - min(req, t) and aft.get(l, 0) == t - min(req, t)
    assert sum(mis.values()) > 0, "configuration must exercise pool exhaustion"
    out = o.sampled()
    assert len(out["tr"]) == sum(smp.values())
    after = o.species()
    taken = before["count"] - after["count"]
    assert np.all(after["count"] <= before["count"])
    key = lambda d: (d["tr"].astype(np.int64) << 40) | (d["start"].astype(np.int64) << 20) | (d["end"] - d["start"])
    u, c = np.unique(key(out), return_counts=True)
    sk = key(before)
    assert np.array_equal(np.sort(sk[taken > 0]), u) and np.array_equal(taken[taken > 0][np.argsort(sk[taken > 0])], c)
    # ids restart per transcript (sampler.go:189,204)
    for t in np.unique(out["tr"]):
        ids = out["id"][out["tr"] == t]
        assert np.array_equal(ids, np.arange(len(ids), dtype=np.uint64))


def test_selection_is_uniform_over_copies(oracle):
    """Multivariate hypergeometric: with one length class of species with known counts, the expected number of
    copies drawn from a species is n * c_s / T; chi-square over 200 sampling seeds."""
    args0 = ["-n", "3000", "-c", "3", "-d", "1:n:(150,1,150,150)", "-e", "0.5", "-m", "0.3", "-si", "2"]
    acc, exp = None, None
    for ss in range(1, 41):
        args = util.parse(args0 + ["-ss", str(ss)])
        names, seqs, levels = util.transcripts(args)
        o = util.run_oracle(oracle, args, names, seqs, levels, oracle.MODE_PHILOX, stages="pcr")
        before = o.species()
        o.sample()
        taken = (before["count"] - o.species()["count"]).astype(np.float64)
        T, n = before["count"].sum(), taken.sum()
        acc = taken if acc is None else acc + taken
        exp = n * before["count"] / T if exp is None else exp + n * before["count"] / T
    # group species into 30 bins of equal expected mass
    order = np.argsort(exp)
    groups = np.array_split(order, 30)
    obs = np.array([acc[g].sum() for g in groups])
    ex = np.array([exp[g].sum() for g in groups])
    assert stats.chisquare(obs, ex * obs.sum() / ex.sum())[1] > 0.001


def test_strand_bias(oracle):
    for bias in (0.0, 1.0, 0.2):
        args = util.parse(["-n", "20000", "-b", str(bias), "-si", "4"])
        names, seqs, levels = util.transcripts(args)
        o = util.run_oracle(oracle, args, names, seqs, levels, oracle.MODE_PHILOX)
        frac = o.sampled()["minus"].mean()
        assert abs(frac - bias) < 4 * math.sqrt(max(bias * (1 - bias), 1e-12) / 20000) + 1e-12


def test_order_free_streams(oracle):
    """Re-ordering / sharding the transcripts must not change any transcript's fragments: emulate two shards."""
    args = util.parse(["-n", "5000", "-si", "8"])
    names, seqs, levels = util.transcripts(args)
    whole = util.run_oracle(oracle, args, names, seqs, levels, oracle.MODE_PHILOX, stages="pcr").species()
    # second shard alone, keyed by its global index through a dummy zero-expression prefix
    part = util.run_oracle(oracle, args, names, [b"A" * 10] * 2 + seqs[2:], [0, 0] + levels[2:], oracle.MODE_PHILOX, stages="pcr").species()
    m = whole["tr"] >= 2
    for k in ("tr", "start", "end", "count0", "count"):
        assert np.array_equal(whole[k][m], part[k]), k


# ---------------------------------------------------------------- the specification is frozen
def _spec_golden():
    import json
    import os
    return json.load(open(os.path.join(util.GOLDEN, "philox_spec.json")))


@pytest.mark.parametrize("name", sorted(_spec_golden()))
def test_philox_spec_is_frozen(oracle, name):
    """tests/golden/philox_spec.json: digests of what the normative restatement produces for every pinned configuration.
    The oracle and the CUDA path are compared with each other everywhere else; this tripwire catches a change that moves
    both together (regenerate with tests/golden/make_spec_golden.py only when the specification changes on purpose)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("make_spec_golden", os.path.join(util.GOLDEN, "make_spec_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = _spec_golden()[name]
    args = util.parse(list(g["argv"]) + ["-si", str(g["seed"])])
    names, seqs, levels = util.transcripts(args)
    o = util.run_oracle(oracle, args, names, seqs, levels, oracle.MODE_PHILOX, stages="pcr")
    sp = o.species()
    o.sample()
    got = mod.digest(sp, o, names)
    assert got == {k: g[k] for k in got}
