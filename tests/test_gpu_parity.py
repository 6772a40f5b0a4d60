"""GPU parity: the CUDA path (through the C ABI) must reproduce the oracle's philox mode bit for bit."""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def h0(gpu):
    return gpu.Handle(0)


def test_philox_blocks(gpu, oracle, h0):
    rng = np.random.default_rng(0)
    ck = rng.integers(0, 2**32, size=(4096, 6), dtype=np.uint64).astype(np.uint32)
    ck[0] = 0
    ck[1] = 0xFFFFFFFF
    ck[2] = [0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0]
    got = h0.probe_philox(ck)
    assert list(got[0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]  # Random123 known answers
    assert list(got[1]) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert list(got[2]) == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    L = oracle.lib()
    out = np.zeros(4, dtype=np.uint32)
    for i in range(0, 4096, 37):
        c, k = np.ascontiguousarray(ck[i, :4]), np.ascontiguousarray(ck[i, 4:])
        L.orc_philox(c.ctypes.data, k.ctypes.data, out.ctypes.data)
        assert list(out) == list(got[i])


@pytest.mark.parametrize("fn,name", [(0, "orc_det_log"), (1, "orc_det_exp"), (2, "orc_det_ndtri"), (3, "orc_det_logfact")])
def test_detmath_bit_exact(oracle, h0, fn, name):
    rng = np.random.default_rng(fn)
    if fn == 0:
        x = np.concatenate([np.exp(rng.uniform(-700, 700, 20000)), rng.uniform(0, 2, 20000), [1.0, 5e-324, 1e308]])
    elif fn == 1:
        x = np.concatenate([rng.uniform(-745, 709, 20000), rng.uniform(-1, 1, 20000), [0.0, -1e-10, 800.0, -800.0]])
    elif fn == 2:
        x = np.concatenate([rng.uniform(0, 1, 30000), 10.0 ** rng.uniform(-300, -1, 5000), 1 - 10.0 ** rng.uniform(-15, -1, 5000)])
        x = x[(x > 0) & (x < 1)]
    else:
        x = np.concatenate([np.arange(0, 300, dtype=np.float64), np.floor(10.0 ** rng.uniform(2, 18, 5000))])
    got = h0.probe_math(fn, x)
    f = getattr(oracle.lib(), name)
    want = np.array([f(float(v)) for v in x])
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))


def test_go_pow_bit_exact(oracle, h0):
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(0, 1, 5000), rng.uniform(1, 5000, 5000)])
    y = rng.choice([-0.1, 0.1, 0.4, 1.0, 2.0, 3.5, -2.3, 8.0, 0.5, -0.5, 0.0], size=len(x))
    got = h0.probe_math(4, x, y)
    want = np.array([oracle.lib().orc_det_go_pow(float(a), float(b)) for a, b in zip(x, y)])
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))


def test_amplify_bit_exact(oracle, h0):
    """thermocycler.go:133-147 on arbitrary (species, count, efficiency): covers the inversion and BTRS branches,
    p in {0, 1}, p > 0.5 and counts up to 1e12."""
    rng = np.random.default_rng(11)
    n = 3000
    tse = rng.integers(0, 2**31, size=(n, 3)).astype(np.uint32)
    n0 = np.concatenate([rng.integers(1, 30, 1000), rng.integers(30, 10**6, 1000), rng.integers(10**6, 10**12, 1000)]).astype(np.uint64)
    p = rng.uniform(0, 1, n)
    p[:10] = 0.0
    p[10:20] = 1.0
    p[20:40] = 1e-9
    for cycles in (1, 11):
        got = h0.probe_amplify(77, tse, n0, p, cycles)
        L = oracle.lib()
        want = np.array([L.orc_px_binomial(77, int(a), int(b), int(c), int(m), float(q), cycles)
                         for (a, b, c), m, q in zip(tse, n0, p)], dtype=np.uint64)
        assert np.array_equal(got, want)


def _compare(oracle, native, argv, seqs_override=None, seed=("-si", "42"), check_fasta=True):
    args = util.parse(list(argv) + list(seed))
    names, seqs, levels = util.transcripts(args) if seqs_override is None else seqs_override
    o = util.run_oracle(oracle, args, names, seqs, levels, oracle.MODE_PHILOX, stages="pcr")
    so = o.species()  # amplified counts, before the draw decrements them
    o.sample()
    h = util.run_gpu(native, args, names, seqs, levels)
    st = h.stats()
    assert (st.low, st.high, st.polya_max) == (o.low, o.high, o.polya_max)
    for kind in (oracle.H_TARGET, oracle.H_POLYA, oracle.H_AFTER_FRAG, oracle.H_AFTER_PCR, oracle.H_SAMPLED,
                 oracle.H_MISSING, oracle.H_AFTER_SAMPLING):
        assert h.hist_dict(kind, 1 << 16) == o.hist_dict(kind), "histogram %d" % kind
    sg = h.species()
    assert st.n_species == len(so["tr"])
    for k in ("tr", "start", "end", "gc"):
        assert np.array_equal(so[k], sg[k]), k
    assert np.array_equal(so["count0"], sg["count0"].astype(np.uint64))
    assert np.array_equal(so["count"], sg["count"]), "amplified copy numbers"
    assert np.array_equal(so["eff"].view(np.uint64), sg["eff"].view(np.uint64)), "per-species efficiency"
    assert st.pool_total == o.pool_total
    fo, fg = o.sampled(), h.sampled()
    for k in ("tr", "start", "end", "minus", "id"):
        assert np.array_equal(fo[k], fg[k]), k
    if check_fasta:
        assert h.fasta(names) == o.fasta(names)
    return o, h


@pytest.mark.parametrize("name", sorted(util.PIN_CONFIGS))
def test_full_path_bit_exact(oracle, gpu, name):
    _compare(oracle, gpu, util.PIN_CONFIGS[name], seed=("-si", str(util.pin_seed(name))))


@pytest.mark.parametrize("name", sorted(util.PIN_CONFIGS))
def test_frozen_spec_digests(gpu, name):
    """The CUDA path against the frozen digests of the specification (tests/golden/philox_spec.json) - no oracle involved."""
    import hashlib
    import json
    import os
    g = json.load(open(os.path.join(util.GOLDEN, "philox_spec.json")))[name]
    args = util.parse(list(g["argv"]) + ["-si", str(g["seed"])])
    names, seqs, levels = util.transcripts(args)
    h = util.run_gpu(gpu, args, names, seqs, levels)
    sp, sm = h.species(), h.sampled()
    dg = lambda *arrs: hashlib.sha256(b"".join(np.ascontiguousarray(a).astype("<u8").tobytes() for a in arrs)).hexdigest()
    st = h.stats()
    assert (int(st.n_species), int(st.n_sampled), int(st.pool_total)) == (g["n_species"], g["n_sampled"], g["pool_total"])
    assert dg(sp["tr"], sp["start"], sp["end"], sp["count0"]) == g["species"] and dg(sp["count"]) == g["counts"]
    assert dg(sm["tr"], sm["start"], sm["end"], sm["minus"], sm["id"]) == g["sampled"]
    assert hashlib.sha256(h.fasta(names)).hexdigest() == g["fasta"]


def test_seed_policy(oracle, gpu):
    """main.go:129-149: -sp / -ss re-key PCR and sampling only."""
    _compare(oracle, gpu, util.C1 + ["-sp", "5", "-ss", "9"])
    _compare(oracle, gpu, util.C1 + ["-sp", "5"])


def test_raw_params(oracle, gpu):
    """-j raw_params.json: raw target + raw GC table, 12 cycles (src/Makefile `t` target)."""
    _compare(oracle, gpu, ["-j", util.RAW_PARAMS, "-n", "20000"])
    _compare(oracle, gpu, ["-j", util.RAW_PARAMS, "-n", "5000", "-jm", "0.5", "-f", "prim_jump"])


def test_fixed_efficiency_doubling(gpu):
    """-e 1.0 -c 11: Binomial(n, 1) = n on every branch => after-PCR histogram = 2048 x after-fragmentation
    (the release binary shows the same identity)."""
    args = util.parse(util.C1 + ["-e", "1.0", "-si", "3"])
    names, seqs, levels = util.transcripts(args)
    h = util.run_gpu(gpu, args, names, seqs, levels, stages="pcr")
    af, ap = h.hist_dict(gpu.H_AFTER_FRAG, 400), h.hist_dict(gpu.H_AFTER_PCR, 400)
    assert ap == {k: v * 2048 for k, v in af.items()}


def test_non_acgt_and_edge_transcripts(oracle, gpu):
    """IUPAC letters (Go map misses / 'N' in RevComp), a transcript shorter than k, zero expression, one molecule."""
    names, seqs, levels = util.synth_transcriptome(40, seed=3, mean_len=900, hi=4000, total_molecules=20000, non_acgt=0.01)
    seqs[0] = b"ACG"
    seqs[1] = b""
    levels[2] = 0
    levels[3] = 1
    _compare(oracle, gpu, ["-n", "30000", "-c", "8"], seqs_override=(names, seqs, levels))
    _compare(oracle, gpu, ["-n", "3000", "-f", "prim_jump:200"], seqs_override=(names, seqs, levels))


@pytest.mark.parametrize("prefix_long", [False, True])
def test_long_molecule_path(oracle, gpu, monkeypatch, prefix_long):
    """A 60 kb transcript with the default -d asks for ~150 breakpoints: exercises k_fragment_long; with RLSIM_PREFIX_LONG
    (opt-in) the prefix sums of such transcripts are built by whole CTAs (k_prefix_long) instead of one warp."""
    if prefix_long:
        monkeypatch.setenv("RLSIM_PREFIX_LONG", "1")
    names, seqs, levels = util.synth_transcriptome(6, seed=8, mean_len=60000, sigma=0.05, lo=50000, hi=70000, total_molecules=600)
    o, h = _compare(oracle, gpu, ["-n", "20000", "-c", "5"], seqs_override=(names, seqs, levels))
    assert int(h.stats().flags) & gpu.F_EAGER  # chunk pipeline: the deferred molecules are served in the chunk's back half
    _compare(oracle, gpu, ["-n", "2000", "-f", "pre_prim:20000", "-d", "1:n:(150,20,50,400)"], seqs_override=(names, seqs, levels))
    # letters outside ACGT in long transcripts: their reverse profiles are explicit arrays, summed by whole CTAs as well
    dirty = list(seqs)
    for t in (1, 4):
        s_ = bytearray(dirty[t])
        s_[20000:20040] = b"N" * 40
        s_[333] = ord("R")
        dirty[t] = bytes(s_)
    _compare(oracle, gpu, ["-n", "20000", "-c", "5", "-f", "after_prim"], seqs_override=(names, dirty, levels))


def test_pool_exhaustion_and_complement(oracle, gpu):
    """C2-like: classes that are drained entirely (missing fragments) and classes where more than half is drawn
    (complement rule of spec 4.6)."""
    o, h = _compare(oracle, gpu, ["-n", "60000", "-c", "2", "-d", "1:n:(150,3,140,160)"], seed=("-si", "5"))
    assert sum(o.hist_dict(oracle.H_MISSING).values()) > 0
    st = h.stats()
    assert st.n_sampled + sum(h.hist_dict(gpu.H_MISSING, 400).values()) == 60000


def test_sharding_invariance(oracle, gpu):
    """Two contiguous shards on two handles with the all-gathered totals reproduce the single-handle result
    (what N GPUs do)."""
    from rlsim_b200 import shard
    args = util.parse(["-n", "40000", "-c", "6", "-si", "13"])
    names, seqs, levels = util.synth_transcriptome(60, seed=4, mean_len=1200, hi=5000, total_molecules=30000)
    whole = util.run_gpu(gpu, args, names, seqs, levels)
    ref = whole.sampled()
    b = shard.contiguous_partition(shard.work_estimate([len(s) for s in seqs], levels), 2)
    parts = [util.run_gpu(gpu, args, names[b[r]:b[r + 1]], seqs[b[r]:b[r + 1]], levels[b[r]:b[r + 1]], stages="pcr", first=b[r])
             for r in range(2)]
    tot = [p.class_totals() for p in parts]
    glob = tot[0] + tot[1]
    outs = []
    for r, p in enumerate(parts):
        p.sample(glob, np.zeros_like(glob) if r == 0 else tot[0])
        outs.append(p.sampled())
    for k in ("tr", "start", "end", "minus", "id"):
        assert np.array_equal(np.concatenate([outs[0][k], outs[1][k]]), ref[k]), k


def test_class_partitioned_selection(gpu):
    """The N-GPU form without redundant work: classes drawn by rank l % N, drawn copy ranks routed to their owners
    (the all-to-all is emulated with device copies between two handles on one GPU).  Identical to one GPU."""
    import torch
    from rlsim_b200 import shard
    args = util.parse(["-n", "40000", "-c", "6", "-si", "13"])
    names, seqs, levels = util.synth_transcriptome(60, seed=4, mean_len=1200, hi=5000, total_molecules=30000)
    whole = util.run_gpu(gpu, args, names, seqs, levels)
    ref = whole.sampled()
    for world in (2, 3):
        b = shard.contiguous_partition(shard.work_estimate([len(s) for s in seqs], levels), world)
        parts = [util.run_gpu(gpu, args, names[b[r]:b[r + 1]], seqs[b[r]:b[r + 1]], levels[b[r]:b[r + 1]], stages="pcr", first=b[r])
                 for r in range(world)]
        all_tot = np.stack([p.class_totals() for p in parts])
        sends, counts = [], []
        for r, p in enumerate(parts):
            c, needed = p.select_classes(all_tot, r)  # size query
            buf = torch.empty((max(needed, 1), 2), dtype=torch.int64, device="cuda:0")
            c, needed2 = p.select_classes(all_tot, r, buf.data_ptr(), buf.shape[0])
            assert needed2 == needed and int(c.sum()) == needed
            sends.append(buf[:needed])
            counts.append([int(x) for x in c])
        torch.cuda.synchronize()
        outs = []
        for r, p in enumerate(parts):
            segs = []
            for s_ in range(world):
                o = sum(counts[s_][:r])
                segs.append(sends[s_][o:o + counts[s_][r]])
            recv = torch.cat(segs) if segs else torch.empty((0, 2), dtype=torch.int64, device="cuda:0")
            torch.cuda.synchronize()
            p.apply_selected(recv.data_ptr() if recv.shape[0] else None, recv.shape[0])
            p.finish_sample()
            outs.append(p.sampled())
        for k in ("tr", "start", "end", "minus", "id"):
            assert np.array_equal(np.concatenate([o[k] for o in outs]), ref[k]), (world, k)
        for kind in (gpu.H_SAMPLED, gpu.H_MISSING, gpu.H_AFTER_SAMPLING):
            assert parts[0].hist_dict(kind, 4096) == whole.hist_dict(kind, 4096)


def test_target_sampling_partition(gpu):
    """a19 on N ranks: each draws a slice of the draw indices; the summed histogram equals the single-rank one."""
    for argv in (["-n", "100000", "-si", "3"], ["-j", util.RAW_PARAMS, "-n", "50000", "-si", "4"]):
        args = util.parse(argv)
        whole = gpu.Handle(0)
        whole.set_model(args.to_params())
        if args.raw_len_probs is not None:
            whole.set_raw_target(args.raw_len_probs.lengths, args.raw_len_probs.probs)
        whole.sample_target(args.req_frags)
        ref = whole.hist(gpu.H_TARGET, 4096)
        acc = np.zeros_like(ref)
        n = args.req_frags
        for r in range(3):
            h = gpu.Handle(0)
            h.set_model(args.to_params())
            if args.raw_len_probs is not None:
                h.set_raw_target(args.raw_len_probs.lengths, args.raw_len_probs.probs)
            a, b = n * r // 3, n * (r + 1) // 3
            h.sample_target_range(a, b - a)
            acc += h.hist(gpu.H_TARGET, 4096)
        assert np.array_equal(acc, ref)
        nz = np.nonzero(acc)[0]
        h.set_target_hist(nz.astype(np.uint32), acc[nz])
        st, sw = h.stats(), whole.stats()
        assert (st.low, st.high, st.n_requested) == (sw.low, sw.high, sw.n_requested)


def test_batched_add_transcripts_equals_single_call(gpu):
    """The Go shim feeds rlsim_add_transcripts in batches: device arrays grow keeping earlier
    batches, the result is that of one call."""
    args = util.parse(["-n", "30000", "-c", "7", "-si", "21"])
    names, seqs, levels = util.synth_transcriptome(90, seed=12, mean_len=1500, hi=6000, total_molecules=25000, non_acgt=0.002)
    one = util.run_gpu(gpu, args, names, seqs, levels)
    h = gpu.Handle(0)
    h.set_model(args.to_params())
    h.sample_target(args.req_frags)
    for a, b in ((0, 1), (1, 37), (37, 38), (38, 90)):
        h.add_transcripts(seqs[a:b], levels[a:b])
    h.build_library()
    h.sample()
    for k in ("tr", "start", "end", "count0", "count"):
        assert np.array_equal(one.species()[k], h.species()[k]), k
    for k in ("tr", "start", "end", "minus", "id"):
        assert np.array_equal(one.sampled()[k], h.sampled()[k]), k
    assert one.fasta(names) == h.fasta(names)
    # a second library on the same handle (buffers reused) gives the same answer again
    h.clear_transcripts()
    h.set_model(args.to_params())
    h.sample_target(args.req_frags)
    h.add_transcripts(seqs, levels)
    h.build_library()
    h.sample()
    assert one.fasta(names) == h.fasta(names)


@pytest.mark.parametrize("flags", [["-f", "after_prim_double"], ["-f", "after_noprim"], ["-f", "pre_prim"], ["-f", "prim_jump"],
                                   ["-f", "after_prim", "-b", "0.3", "-flg", "0.2"]])
def test_eager_chunk_pipeline_equals_monolithic(gpu, monkeypatch, flags):
    """rlsim_add_transcripts fragments / deduplicates / amplifies each copied chunk while later chunks are in flight
    (after_* methods); the result is bit-identical to the monolithic rlsim_fragment + rlsim_pcr over resident
    transcripts, for any chunking, for repeated build_library calls, and for batches added in several calls."""
    args = util.parse(["-n", "40000", "-c", "9", "-si", "5", "-eg", "(1.0,0.5,0.8)"] + flags)
    names, seqs, levels = util.synth_transcriptome(400, seed=31, mean_len=1200, hi=5000, total_molecules=60000, non_acgt=0.001)
    levels = list(levels)
    levels[3] = levels[4] = 0

    def run(first_builds=1, batches=((0, 400),)):
        h = gpu.Handle(0)
        h.set_model(args.to_params())
        h.sample_target(args.req_frags)
        h.set_first_index(7)
        for a, b in batches:
            h.add_transcripts(seqs[a:b], levels[a:b])
        for _ in range(first_builds):
            h.build_library()
        h.sample()
        return h

    monkeypatch.setenv("RLSIM_NO_EAGER", "1")
    mono = run()
    monkeypatch.delenv("RLSIM_NO_EAGER")
    variants = []
    for chunk in ("4096", "50000", "100000000"):
        monkeypatch.setenv("RLSIM_CHUNK_BYTES", chunk)
        variants.append(run())
    monkeypatch.setenv("RLSIM_CHUNK_BYTES", "30000")
    variants.append(run(first_builds=2))                                   # second build recomputes from resident data
    variants.append(run(batches=((0, 1), (1, 150), (150, 151), (151, 400))))  # eager across several add calls
    # the grouped form of the pipeline (host round trips per group; what the asynchronous default replaced)
    monkeypatch.setenv("RLSIM_EAGER_SYNC", "1")
    variants.append(run())
    variants.append(run(batches=((0, 150), (150, 400))))
    monkeypatch.delenv("RLSIM_EAGER_SYNC")
    # the asynchronous form without stream priorities
    monkeypatch.setenv("RLSIM_NO_PRIORITY", "1")
    variants.append(run())
    monkeypatch.delenv("RLSIM_NO_PRIORITY")
    # a buffer estimate that turns out too small in the first group, or in a later one (two add calls make at least two
    # groups whatever the copy timing): the monolithic path takes over and recomputes from the resident transcripts
    for fail_at, batches in (("0", ((0, 400),)), ("1", ((0, 200), (200, 400)))):
        monkeypatch.setenv("RLSIM_EAGER_TEST_FAIL", fail_at)
        h = run(batches=batches)
        assert not (int(h.stats().flags) & gpu.F_EAGER), fail_at
        variants.append(h)
    monkeypatch.delenv("RLSIM_EAGER_TEST_FAIL")
    # the eager pipeline serves the after_* methods (pre_prim / prim_jump always take the monolithic path)
    assert bool(int(variants[0].stats().flags) & gpu.F_EAGER) == flags[1].startswith("after")
    assert not (int(mono.stats().flags) & gpu.F_EAGER)
    for h in variants:
        for k in ("tr", "start", "end", "count0", "count"):
            assert np.array_equal(mono.species()[k], h.species()[k]), k
        for k in ("tr", "start", "end", "minus", "id"):
            assert np.array_equal(mono.sampled()[k], h.sampled()[k]), k
        for which in (gpu.H_AFTER_FRAG, gpu.H_POLYA, gpu.H_AFTER_PCR):
            assert mono.hist_dict(which, 6000) == h.hist_dict(which, 6000)
        assert mono.fasta(names) == h.fasta(names)
        sm, sh = mono.stats(), h.stats()
        assert (sm.n_fragments, sm.n_species, sm.pool_total, sm.n_sampled) == (sh.n_fragments, sh.n_species, sh.pool_total, sh.n_sampled)


def test_mirrored_reverse_profile_equals_explicit_arrays(oracle, gpu, monkeypatch):
    """The quantised k-mer table is reverse-complement symmetric (checked on the device when it is built), so
    transcripts without non-ACGT letters carry no reverse prefix sums: the reverse priming draw searches the forward
    arrays mirrored.  Same fragments as with explicit reverse arrays (RLSIM_NO_MIRROR) and as the oracle."""
    args = util.parse(["-n", "30000", "-c", "7", "-si", "77", "-f", "after_prim_double"])
    names, seqs, levels = util.synth_transcriptome(120, seed=5, mean_len=1400, hi=6000, total_molecules=40000, non_acgt=0.0)
    rng = np.random.default_rng(3)
    seqs = list(seqs)
    for t in (0, 7, 8, 50, 119):  # a few transcripts with letters outside ACGT keep their explicit reverse arrays
        b = bytearray(seqs[t])
        for pos in rng.integers(0, len(b), size=5):
            b[pos] = ord("N") if pos % 2 else ord("R")
        seqs[t] = bytes(b)
    seqs[20] = seqs[20][:6]   # exactly k bases + tail
    seqs[21] = seqs[21][:3]
    mirrored = util.run_gpu(gpu, args, names, seqs, levels)
    monkeypatch.setenv("RLSIM_NO_MIRROR", "1")
    explicit = util.run_gpu(gpu, args, names, seqs, levels)
    monkeypatch.delenv("RLSIM_NO_MIRROR")
    for k in ("tr", "start", "end", "count0", "count"):
        assert np.array_equal(mirrored.species()[k], explicit.species()[k]), k
    assert mirrored.fasta(names) == explicit.fasta(names)
    o = util.run_oracle(oracle, args, names, seqs, levels, oracle.MODE_PHILOX)
    assert mirrored.fasta(names) == o.fasta(names)
    for t in (0, 1, 7, 20, 21, 119):  # reverse prefix sums reconstructed from either representation
        n = len(seqs[t]) + int(mirrored.stats().polya_max) + 1
        assert np.array_equal(mirrored.probe_prefix(t, 1, n), explicit.probe_prefix(t, 1, n)), t


def test_degenerate_inputs(oracle, gpu):
    """Nothing expressed; a single molecule; primer length 1 and 12 (LUT in global memory); fixed-length tails."""
    names, seqs, levels = util.synth_transcriptome(5, seed=2, mean_len=800, hi=2000, total_molecules=100)
    args = util.parse(["-n", "100", "-si", "1"])
    h = util.run_gpu(gpu, args, names, seqs, [0] * 5)
    st = h.stats()
    assert (st.n_molecules, st.n_fragments, st.n_species, st.n_sampled, st.pool_total) == (0, 0, 0, 0, 0)
    assert sum(h.hist_dict(gpu.H_MISSING, 400).values()) == 100 and h.fasta(names) == b""
    _compare(oracle, gpu, ["-n", "50"], seqs_override=(names, seqs, [1, 0, 0, 0, 0]))
    _compare(oracle, gpu, ["-n", "2000", "-k", "1"], seqs_override=(names, seqs, levels))
    _compare(oracle, gpu, ["-n", "2000", "-k", "12", "-p", "9.5"], seqs_override=(names, seqs, levels))
    _compare(oracle, gpu, ["-n", "2000", "-a", "1:n:(40,5,33,33)", "-d", "1:n:(120,10,120,120)"], seqs_override=(names, seqs, levels))


def test_amplify_large_sample_bit_exact(oracle, h0):
    """400 000 species x 15 cycles across all count regimes: the FP32 pre-decision of the BTRS acceptance test must
    never disagree with the exact FP64 test of the spec (a single wrong decision changes a copy number)."""
    rng = np.random.default_rng(2024)
    n = 400_000
    tse = rng.integers(0, 2**31, size=(n, 3)).astype(np.uint32)
    n0 = np.concatenate([rng.integers(1, 4, n // 2), rng.integers(20, 5000, n // 4),
                         np.floor(10.0 ** rng.uniform(4, 12.5, n - n // 2 - n // 4))]).astype(np.uint64)
    p = np.concatenate([rng.uniform(0.05, 0.95, n - 1000), rng.uniform(0.4999, 0.5001, 500), rng.uniform(0, 0.01, 500)])
    got = h0.probe_amplify(5, tse, n0, p, 15)
    L = oracle.lib()
    idx = np.concatenate([np.arange(0, n, 7), np.arange(n - 1000, n)])  # oracle on a 1/7 subsample + all edge probabilities
    want = np.array([L.orc_px_binomial(5, int(tse[i, 0]), int(tse[i, 1]), int(tse[i, 2]), int(n0[i]), float(p[i]), 15) for i in idx],
                    dtype=np.uint64)
    assert np.array_equal(got[idx], want)
