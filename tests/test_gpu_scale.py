"""GPU parity at the shapes BASELINE.json names, and direct probes of the deterministic building blocks.

* the benchmarked path itself: a C3-shaped set pushed through rlsim_add_transcripts with the default 32 MB chunks (several
  eager groups), bit for bit against the oracle's philox mode - species, copy numbers, histograms, records, FASTA text;
* C5 at its stated shape (C1 input, -m 100, 30 cycles: 3.56 M species, copy numbers up to 1e11);
* device tables against the oracle's restatement: K for all 4^k k-mers (nnthermo.go:126-175; bitwise, and within the 1e-6
  relative tolerance north_star states against libm), priming prefix sums forward and reverse incl. non-ACGT windows
  (nnthermo.go:177-201), GC prefix counts (thermocycler.go:158-166);
* the host link: packed records == unpacked records, pageable caller buffers (staged) == pinned ones.
"""
import ctypes
import hashlib

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu

C3_FLAGS = ["-d", "0.9:sn:(600,50,4,300,2000) + 0.1:n:(700,1,600,2000)", "-f", "after_prim_double", "-c", "15",
            "-eg", "(1.0,0.5,0.8)", "-el", "(0.1,0.7,1.0)"]


def _species_equal(so, sg):
    for k in ("tr", "start", "end", "gc"):
        assert np.array_equal(so[k], sg[k]), k
    assert np.array_equal(so["count0"], sg["count0"].astype(np.uint64))
    assert np.array_equal(so["count"], sg["count"]), "amplified copy numbers"
    assert np.array_equal(so["eff"].view(np.uint64), sg["eff"].view(np.uint64)), "per-species efficiency"


def test_c3_shaped_multichunk_eager_bit_exact(oracle, gpu, monkeypatch):
    """What bench.py times: 30 000 C3-shaped transcripts (67 MB of bases: three 32 MB chunks, at least two eager groups),
    1.5 M molecules, C3 flags, through add_transcripts_raw - against the oracle."""
    from rlsim_b200 import synth
    n_tr, n_frags = 30_000, 1_500_000
    bases, off, levels = synth.transcriptome(n_tr, seed=20261019, total_molecules=1.5e6)
    args = util.parse(["-n", str(n_frags), "-si", "42"] + C3_FLAGS)
    seqs = [bases[int(off[i]):int(off[i + 1])].tobytes() for i in range(n_tr)]
    names = [b"T%07d" % i for i in range(n_tr)]
    o = util.run_oracle(oracle, args, names, seqs, [int(x) for x in levels], oracle.MODE_PHILOX, stages="pcr")
    so = o.species()
    o.sample()

    h = gpu.Handle(0)
    h.set_model(args.to_params())
    h.sample_target(args.req_frags)
    h.set_first_index(0)
    h.add_transcripts_raw(bases, off, levels, n_tr)  # pageable numpy buffers: staged through pinned memory
    h.build_library()
    h.sample()
    st = h.stats()
    flags = int(st.flags)
    assert flags & gpu.F_EAGER, "the eager chunk pipeline did not run"
    assert (flags >> gpu.F_GROUPS_SHIFT) & 0xff >= 2, "expected at least two eager groups at the default chunk size"
    assert flags & gpu.F_PAGEABLE_STAGED
    assert (st.low, st.high, st.polya_max) == (o.low, o.high, o.polya_max)
    for kind in range(7):
        assert h.hist_dict(kind, 1 << 16) == o.hist_dict(kind), "histogram %d" % kind
    assert st.n_species == len(so["tr"]) and st.pool_total == o.pool_total
    _species_equal(so, h.species())
    fo, fg = o.sampled(), h.sampled()
    for k in ("tr", "start", "end", "minus", "id"):
        assert np.array_equal(fo[k], fg[k]), k
    assert hashlib.sha256(h.fasta(names)).digest() == hashlib.sha256(o.fasta(names)).digest()

    # the same library from pinned host memory (direct DMA) and on the monolithic path: identical
    import torch
    pb, po, pl = (torch.from_numpy(x.view(np.int64) if x.dtype == np.uint64 else x).pin_memory() for x in (bases, off, levels))
    h2 = gpu.Handle(0)
    h2.set_model(args.to_params())
    h2.sample_target(args.req_frags)
    h2.set_first_index(0)
    h2.add_transcripts_raw(pb, po, pl, n_tr)
    assert not (int(h2.stats().flags) & gpu.F_PAGEABLE_STAGED)
    h2.build_library()
    h2.build_library()  # second call: monolithic recomputation from the resident transcripts
    assert not (int(h2.stats().flags) & gpu.F_EAGER)
    h2.sample()
    _species_equal(so, h2.species())
    f2 = h2.sampled()
    for k in ("tr", "start", "end", "minus", "id"):
        assert np.array_equal(fo[k], f2[k]), k

    # resident transcripts through the grouped pipeline (what bench.py's device-resident step runs): PCR of one group of
    # transcripts overlaps the fragmentation of the next
    h3 = gpu.Handle(0)
    h3.set_model(args.to_params())
    h3.sample_target(args.req_frags)
    h3.set_first_index(0)
    monkeypatch.setenv("RLSIM_NO_EAGER", "1")
    h3.add_transcripts_raw(pb, po, pl, n_tr)
    monkeypatch.delenv("RLSIM_NO_EAGER")
    monkeypatch.setenv("RLSIM_GROUPS", "3")
    h3.build_library()
    monkeypatch.delenv("RLSIM_GROUPS")
    f3 = int(h3.stats().flags)
    assert f3 & gpu.F_EAGER and (f3 >> gpu.F_GROUPS_SHIFT) & 0xff == 3
    h3.sample()
    _species_equal(so, h3.species())
    r3 = h3.sampled()
    for k in ("tr", "start", "end", "minus", "id"):
        assert np.array_equal(fo[k], r3[k]), k
    for kind in range(7):
        assert h3.hist_dict(kind, 1 << 16) == o.hist_dict(kind), "histogram %d" % kind


def test_c5_stated_shape_bit_exact(oracle, gpu):
    """BASELINE.json configs[4]: C1 input at -m 100 (2.36 M molecules), 30 cycles, -eg (1.0,0.5,0.8), 1e6 fragments:
    3.56 M fragments, pool 1.8e13, per-length totals beyond 2^32 - the large-count BTRS branch at scale."""
    args = util.parse(["-n", "1000000", "-m", "100", "-c", "30", "-eg", "(1.0,0.5,0.8)", "-si", "42"])
    names, seqs, levels = util.transcripts(args)
    o = util.run_oracle(oracle, args, names, seqs, levels, oracle.MODE_PHILOX, stages="pcr")
    so = o.species()
    o.sample()
    h = util.run_gpu(gpu, args, names, seqs, levels)
    st = h.stats()
    # (SURVEY 8d's "3.56 M species" counts fragments: five transcripts hold ~1.4 M distinct (start, end) pairs)
    assert st.n_molecules > 2_000_000 and st.n_fragments > 3_000_000 and st.n_species > 1_000_000 and st.pool_total > 10**13
    assert max(h.hist_dict(gpu.H_AFTER_PCR, 1 << 12).values()) > 2**32
    for kind in range(7):
        assert h.hist_dict(kind, 1 << 16) == o.hist_dict(kind), "histogram %d" % kind
    assert st.n_species == len(so["tr"]) and st.pool_total == o.pool_total
    _species_equal(so, h.species())
    fo, fg = o.sampled(), h.sampled()
    for k in ("tr", "start", "end", "minus", "id"):
        assert np.array_equal(fo[k], fg[k]), k
    assert hashlib.sha256(h.fasta(names)).digest() == hashlib.sha256(o.fasta(names)).digest()


# ---- direct probes ---------------------------------------------------------------------------------------------------
APPENDIX_C = {  # SURVEY.md Appendix C (T = 5.0, k = 6), derived from nnthermo.go:68-170
    b"AAAAAA": 27.420537338477043, b"TTTTTT": 27.420537338477043, b"GGGGGG": 52.03896441320523,
    b"ACGTAC": 56.17841501711723, b"CGCGCG": 161.4976741063506, b"GAATTC": 45.96815493765271,
    b"ACCCCA": 36.74549747563298, b"GCATGC": 75.75817022252053,
}


def _kmers(k):
    n = 1 << (2 * k)
    c = np.arange(n, dtype=np.uint32)
    cols = [np.frombuffer(b"ACGT", dtype=np.uint8)[(c >> (2 * i)) & 3] for i in range(k)]  # base i in bits [2i, 2i+2)
    return np.stack(cols, axis=1)


@pytest.mark.parametrize("k,T", [(6, 5.0), (8, 2.0), (4, 0.3), (1, 5.0)])
def test_probe_kmer_lut_vs_oracle(oracle, gpu, k, T):
    """a2: K for every ACGT k-mer - bitwise against the oracle's deterministic restatement, within 1e-6 relative of the
    libm evaluation (north_star's tolerance for the affinities), and the fixed-point weights of spec 4.3."""
    args = util.parse(["-n", "10", "-k", str(k), "-p", str(T)])
    h = gpu.Handle(0)
    h.set_model(args.to_params())
    K, w = h.probe_kmer_lut(k)
    km = _kmers(k)
    L = oracle.lib()
    step = 1 if k <= 6 else 13
    idx = np.arange(0, len(K), step)
    want_det = np.array([L.orc_kmer_affinity(km[i].tobytes(), k, T, 1) for i in idx])
    want_libm = np.array([L.orc_kmer_affinity(km[i].tobytes(), k, T, 0) for i in idx])
    assert np.array_equal(K[idx].view(np.uint64), want_det.view(np.uint64))
    assert np.max(np.abs(K[idx] - want_libm) / want_libm) <= 1e-6
    if (k, T) == (6, 5.0):
        by_kmer = {km[i].tobytes(): K[i] for i in range(len(K))}
        for kmer, val in APPENDIX_C.items():
            assert abs(by_kmer[kmer] - val) / val <= 1e-6, kmer
        assert abs(K.min() - 22.45726863508428) / 22.457 <= 1e-6 and abs(K.max() - 161.4976741063506) / 161.5 <= 1e-6
        assert abs(K.sum() - 201711.76817515807) / 201711.768 <= 1e-6
    scale = 67108864.0 / K.max()
    x = K * scale + 0.5
    want_w = np.where(x < 67108863.0, np.maximum(np.floor(x), 1.0), 67108863.0).astype(np.uint32)
    assert np.array_equal(w, want_w)
    assert w.min() >= 1 and w.max() == 67108863
    if T < 1.0:
        assert (want_w == 1).any(), "the low-temperature table should reach the lower clamp"


@pytest.mark.parametrize("variant", ["default", "lane_stores", "long_by_cta"])
def test_probe_prefix_and_gc_vs_oracle(oracle, gpu, monkeypatch, variant):
    """a3 + a12: the priming prefix sums the searches read (forward, explicit reverse, mirrored reverse) against the
    cumulative sums of the oracle's quantised weights - windows with letters outside ACGT included - and the GC prefix
    counts against a byte scan.  Variants: the fine sums stored by the lanes instead of leaving shared memory as bulk
    asynchronous copies (RLSIM_NO_PREFIX_BULK), long transcripts summed by whole CTAs (RLSIM_PREFIX_LONG)."""
    if variant == "lane_stores":
        monkeypatch.setenv("RLSIM_NO_PREFIX_BULK", "1")
    if variant == "long_by_cta":
        monkeypatch.setenv("RLSIM_PREFIX_LONG", "1")
    args = util.parse(["-n", "3000", "-c", "3", "-si", "9", "-f", "after_prim_double"])
    names, seqs, levels = util.synth_transcriptome(12, seed=77, mean_len=900, hi=3000, total_molecules=3000)
    rng = np.random.default_rng(5)
    seqs = list(seqs)
    for t in (1, 4):  # dirty transcripts: explicit reverse arrays
        b = bytearray(seqs[t])
        for pos in rng.integers(0, len(b), size=9):
            b[pos] = ord("NRYK"[pos % 4])
        b[10:18] = b"NNNNNNNN"  # a run longer than k: windows with no valid doublet at all
        seqs[t] = bytes(b)
    seqs[7] = seqs[7][:6]
    seqs[5] = (seqs[5] * 40)[:256 * 3 + 6 + 1]      # profile of exactly three sub-tiles + 1 entry (poly(A) adds to it below)
    seqs[11] = (seqs[11] * 40)[:9000]               # above the threshold of k_prefix_long
    h = util.run_gpu(gpu, args, names, seqs, levels, stages="fragment")
    pa = int(h.stats().polya_max)
    k, T = 6, 5.0
    L = oracle.lib()
    for t in (0, 1, 4, 5, 7, 11):
        plus = seqs[t] + b"A" * pa
        n = len(plus)
        proflen = n - k
        for reverse in (0, 1):
            w = np.zeros(max(proflen, 1), dtype=np.uint64)
            L.orc_quantized_weights(plus, n, k, T, reverse, w.ctypes.data_as(ctypes.c_void_p))
            want = np.concatenate([[0], np.cumsum(w[:proflen])]).astype(np.uint64)
            got = h.probe_prefix(t, reverse, proflen + 1)
            assert np.array_equal(got, want), (t, reverse)
        g = h.probe_gc_prefix(t, n + 1)
        a = np.frombuffer(plus, dtype=np.uint8)
        want_gc = np.concatenate([[0], np.cumsum((a == ord("G")) | (a == ord("C")))]).astype(np.uint32)
        assert np.array_equal(g, want_gc), t
        for x in (0, 1, 31, 32, 33, n // 2, n):
            assert int(g[x]) == L.orc_gc_count(plus, 0, x)


# ---- low -p, runs of N (ADVICE: the 26-bit quantisation must not abort what the reference completes) ----------------
@pytest.mark.parametrize("flags", [["-p", "0.5"], ["-p", "0.3", "-k", "8"], ["-p", "1.1", "-k", "12"]])
def test_low_priming_temperature(oracle, gpu, flags):
    names, seqs, levels = util.synth_transcriptome(20, seed=6, mean_len=900, hi=3000, total_molecules=4000)
    _compare_records(oracle, gpu, ["-n", "3000", "-c", "4", "-si", "3"] + flags, names, seqs, levels)


def test_runs_of_n(oracle, gpu):
    """Runs of N longer than the shortest fragment: every window inside has K ~ 1 against Kmax ~ 1e10 at -p 1.5."""
    names, seqs, levels = util.synth_transcriptome(10, seed=16, mean_len=1500, hi=3000, total_molecules=3000)
    seqs = list(seqs)
    for t in range(0, 10, 2):
        b = bytearray(seqs[t])
        b[100:500] = b"N" * 400
        seqs[t] = bytes(b)
    _compare_records(oracle, gpu, ["-n", "3000", "-c", "4", "-si", "4", "-p", "1.5"], names, seqs, levels)
    _compare_records(oracle, gpu, ["-n", "2000", "-c", "4", "-si", "4", "-p", "1.5", "-f", "prim_jump"], names, seqs, levels)


def _compare_records(oracle, gpu, argv, names, seqs, levels):
    args = util.parse(argv)
    o = util.run_oracle(oracle, args, names, seqs, levels, oracle.MODE_PHILOX)
    h = util.run_gpu(gpu, args, names, seqs, levels)
    assert h.stats().n_fragments > 0
    fo, fg = o.sampled(), h.sampled()
    for k in ("tr", "start", "end", "minus", "id"):
        assert np.array_equal(fo[k], fg[k]), k
    assert h.fasta(names) == o.fasta(names)
    return h


# ---- the host link ------------------------------------------------------------------------------------------------
def test_packed_records_equal_unpacked(gpu):
    """rlsim_get_sampled_packed (6 B per record: per-transcript counts + start + length|strand) expands to exactly what
    rlsim_get_sampled returns (21 B per record), with 2- and 4-byte length elements, and for a shard with first_index."""
    names, seqs, levels = util.synth_transcriptome(50, seed=21, mean_len=1100, hi=4000, total_molecules=20000)
    levels = list(levels)
    levels[0] = levels[17] = levels[49] = 0  # transcripts without records at the edges and inside
    args = util.parse(["-n", "25000", "-c", "6", "-si", "8", "-b", "0.3"])
    h = util.run_gpu(gpu, args, names, seqs, levels, first=1000)
    un = h.sampled_unpacked()
    n = int(h.stats().n_sampled)
    assert n > 10000 and un["minus"].sum() not in (0, n)
    for eb in (2, 4):
        packed, eb_ = h.sampled_packed(elem_bytes=eb)
        assert int(packed["tr_counts"].sum()) == n and packed["tr_counts"][0] == 0 and packed["tr_counts"][17] == 0
        ex = gpu.Handle.unpack_records(packed, eb_, 1000, n)
        for k in ("tr", "start", "end", "minus", "id"):
            assert np.array_equal(ex[k], un[k]), (eb, k)
    # lengths that need the 4-byte form
    # (size selection is per exact length: a narrow target and many molecules so that some requested lengths exist)
    args = util.parse(["-n", "300", "-c", "2", "-si", "8", "-d", "1:n:(40000,5,39980,40020)"])
    names, seqs, levels = util.synth_transcriptome(4, seed=2, mean_len=90000, sigma=0.05, lo=80000, hi=100000, total_molecules=300000)
    h = util.run_gpu(gpu, args, names, seqs, levels)
    with pytest.raises(gpu.RlsimError):
        h.sampled_packed(elem_bytes=2)
    packed, eb = h.sampled_packed()
    assert eb == 4
    ex, un = gpu.Handle.unpack_records(packed, 4, 0, int(h.stats().n_sampled)), h.sampled_unpacked()
    assert int(h.stats().n_sampled) > 0
    for k in ("tr", "start", "end", "minus", "id"):
        assert np.array_equal(ex[k], un[k]), k


def test_pageable_and_pinned_buffers_agree(gpu, monkeypatch):
    """Pageable caller memory is staged through the pinned ring by the copy team (both directions); the bytes that
    arrive are those of the direct DMA path.  Small chunks make the ring wrap many times."""
    import torch
    from rlsim_b200 import synth
    n_tr = 3000
    bases, off, levels = synth.transcriptome(n_tr, seed=99, total_molecules=60000)
    args = util.parse(["-n", "400000", "-si", "5"] + C3_FLAGS)

    def run(b, o, l):
        h = gpu.Handle(0)
        h.set_model(args.to_params())
        h.sample_target(args.req_frags)
        h.add_transcripts_raw(b, o, l, n_tr)
        h.build_library()
        h.sample()
        return h

    monkeypatch.setenv("RLSIM_CHUNK_BYTES", "300000")
    monkeypatch.setenv("RLSIM_STAGE_THREADS", "3")
    staged = run(bases, off, levels)
    assert int(staged.stats().flags) & gpu.F_PAGEABLE_STAGED
    pinned = run(*(torch.from_numpy(x.view(np.int64) if x.dtype == np.uint64 else x).pin_memory() for x in (bases, off, levels)))
    assert not (int(pinned.stats().flags) & gpu.F_PAGEABLE_STAGED)
    monkeypatch.setenv("RLSIM_NO_STAGING", "1")
    driver = run(bases, off, levels)  # the driver's own pageable path
    monkeypatch.delenv("RLSIM_NO_STAGING")
    names = [b"T%d" % i for i in range(n_tr)]
    ref = pinned.sampled_unpacked()
    assert int(pinned.stats().n_sampled) > 300000  # > 1 MB per array: the staged device-to-host path is taken
    for other in (staged, driver):
        for k in ("tr", "start", "end", "count0", "count"):
            assert np.array_equal(pinned.species()[k], other.species()[k]), k
        got = other.sampled_unpacked()  # numpy destinations: staged device-to-host
        for k in ("tr", "start", "end", "minus", "id"):
            assert np.array_equal(ref[k], got[k]), k
    into = dict(tr=torch.empty(len(ref["tr"]), dtype=torch.int32).pin_memory(), start=torch.empty(len(ref["tr"]), dtype=torch.int32).pin_memory(),
                end=torch.empty(len(ref["tr"]), dtype=torch.int32).pin_memory(), minus=torch.empty(len(ref["tr"]), dtype=torch.uint8).pin_memory(),
                id=torch.empty(len(ref["tr"]), dtype=torch.int64).pin_memory())
    staged.sampled_unpacked(into)
    for k in ("tr", "start", "end", "minus", "id"):
        assert np.array_equal(ref[k].astype(np.int64), into[k].numpy().astype(np.int64) & (0xffffffff if k != "id" else -1)), k
    assert staged.fasta(names) == pinned.fasta(names)


def test_packed_input_equals_ascii_input(oracle, gpu, monkeypatch):
    """rlsim_add_transcripts_packed (2-bit codes + N mask, 0.375 B per base on the host link) builds the library
    rlsim_add_transcripts builds from the same sequences: species, records and FASTA text, with runs of N, transcripts of
    every length mod 32 (empty ones included), a batch that starts in the middle of a code byte, two batches on one handle,
    pinned and pageable arrays, many small chunks; and the oracle agrees with both."""
    import torch
    rng = np.random.default_rng(77)
    names, seqs, levels = util.synth_transcriptome(150, seed=31, mean_len=900, hi=3000, total_molecules=40000)
    seqs, levels = list(seqs), list(levels)
    for t in range(0, 150, 7):  # N runs and single Ns
        b = bytearray(seqs[t])
        a = int(rng.integers(0, max(1, len(b) - 60)))
        r = int(rng.integers(1, 50))
        b[a:a + r] = b"N" * len(b[a:a + r])
        seqs[t] = bytes(b)
    for t in range(40):  # every length mod 32, very short ones, an empty one
        seqs[100 + t] = seqs[100 + t][:t]
    argv = ["-n", "30000", "-c", "6", "-si", "12"]
    args = util.parse(argv)
    lead = 5  # the batch starts at base 5 of the caller's arrays: inside a code byte and a mask byte
    bases = np.frombuffer(b"ACGTN" + b"".join(seqs), dtype=np.uint8)
    off = np.zeros(len(seqs) + 1, np.uint64)
    off[1:] = np.cumsum([len(x) for x in seqs])
    off += lead
    lv = np.asarray(levels, dtype=np.uint64)
    codes, nmask = gpu.pack_bases(bases)
    assert nmask is not None and len(codes) == (len(bases) + 3) // 4

    def run(add, split=None):
        h = gpu.Handle(0)
        h.set_model(args.to_params())
        h.sample_target(args.req_frags)
        n = len(seqs)
        if split is None:
            add(h, off, lv, n)
        else:  # two batches
            add(h, off[:split + 1].copy(), lv[:split].copy(), split)
            add(h, off[split:].copy(), lv[split:].copy(), n - split)
        h.build_library()
        h.sample()
        return h

    ascii_h = run(lambda h, o, l, n: h.add_transcripts_raw(bases, o, l, n))
    ref_sp, ref = ascii_h.species(), ascii_h.sampled()
    assert int(ascii_h.stats().n_sampled) > 5000
    pin = lambda x: torch.from_numpy(x).pin_memory()
    variants = {
        "pageable": lambda: run(lambda h, o, l, n: h.add_transcripts_packed(codes, nmask, o, l, n)),
        "pinned": lambda: run(lambda h, o, l, n: h.add_transcripts_packed(pin(codes), pin(nmask), o, l, n)),
        "two_batches": lambda: run(lambda h, o, l, n: h.add_transcripts_packed(codes, nmask, o, l, n), split=61),
    }
    monkeypatch.setenv("RLSIM_CHUNK_BYTES", "7001")
    variants["small_chunks"] = lambda: run(lambda h, o, l, n: h.add_transcripts_packed(pin(codes), pin(nmask), o, l, n))
    for name, make in variants.items():
        if name == "small_chunks":
            monkeypatch.setenv("RLSIM_CHUNK_BYTES", "7001")
        else:
            monkeypatch.delenv("RLSIM_CHUNK_BYTES", raising=False)
        h = make()
        sp, got = h.species(), h.sampled()
        for k in ("tr", "start", "end", "count0", "count"):
            assert np.array_equal(ref_sp[k], sp[k]), (name, k)
        for k in ("tr", "start", "end", "minus", "id"):
            assert np.array_equal(ref[k], got[k]), (name, k)
        assert h.fasta(names) == ascii_h.fasta(names), name
        assert int(h.stats().flags) & gpu.F_EAGER, name
    monkeypatch.delenv("RLSIM_CHUNK_BYTES", raising=False)
    # no mask at all: a clean set
    clean = [bytes(x).replace(b"N", b"A") for x in seqs]
    cb = np.frombuffer(b"".join(clean), dtype=np.uint8)
    o0 = off - lead
    cc, cm = gpu.pack_bases(cb)
    assert cm is None
    a = run(lambda h, o, l, n: h.add_transcripts_raw(cb, o0, l, n))
    b = run(lambda h, o, l, n: h.add_transcripts_packed(cc, None, o0, l, n))
    for k in ("tr", "start", "end", "minus", "id"):
        assert np.array_equal(a.sampled()[k], b.sampled()[k]), k
    assert a.fasta(names) == b.fasta(names)
    # and the oracle on the N-carrying set
    o = util.run_oracle(oracle, args, names, seqs, levels, oracle.MODE_PHILOX)
    fo = o.sampled()
    for k in ("tr", "start", "end", "minus", "id"):
        assert np.array_equal(fo[k], ref[k]), k


# ---- fused prefix + fragmentation kernel (prefix sums in shared memory) --------------------------------------------
def test_fused_path_variants(oracle, gpu, monkeypatch):
    """after_prim / after_prim_double keep the priming prefix sums in shared memory: one warp per transcript with a compact
    profile (k_fused_warp), transcripts with letters outside ACGT through the CTA-cooperative kernel (k_fused, explicit
    reverse sums, sub-batching), deferred molecules through k_fused_long.  Same fragments as the HBM path (RLSIM_NO_FUSED)
    and as the oracle for: all three slice sizes, items split by molecule range (a heavily expressed transcript),
    zero-expression gaps, and a transcript whose sums do not fit (fallback to the HBM path)."""
    monkeypatch.setenv("RLSIM_FUSED", "1")  # the fused kernels are opt-in

    def both(argv, names, seqs, levels, expect_fused=True, eager=True):
        h = _compare_records(oracle, gpu, argv, names, seqs, levels)
        assert bool(int(h.stats().flags) & gpu.F_FUSED) == expect_fused, argv
        monkeypatch.setenv("RLSIM_NO_FUSED", "1")
        h2 = util.run_gpu(gpu, util.parse(argv), names, seqs, levels)
        monkeypatch.delenv("RLSIM_NO_FUSED")
        assert not (int(h2.stats().flags) & gpu.F_FUSED)
        for k in ("tr", "start", "end", "count0", "count"):
            assert np.array_equal(h.species()[k], h2.species()[k]), k
        for kind in (gpu.H_AFTER_FRAG, gpu.H_POLYA):
            assert h.hist_dict(kind, 1 << 16) == h2.hist_dict(kind, 1 << 16)
        monkeypatch.setenv("RLSIM_FUSED_CTA", "1")  # the CTA-cooperative kernel for every transcript (normally: only the redo list)
        h4 = util.run_gpu(gpu, util.parse(argv), names, seqs, levels)
        monkeypatch.delenv("RLSIM_FUSED_CTA")
        assert bool(int(h4.stats().flags) & gpu.F_FUSED) == expect_fused
        for k in ("tr", "start", "end", "count0", "count"):
            assert np.array_equal(h.species()[k], h4.species()[k]), k
        for kind in (gpu.H_AFTER_FRAG, gpu.H_POLYA):
            assert h.hist_dict(kind, 1 << 16) == h4.hist_dict(kind, 1 << 16)
        if not eager:
            return
        monkeypatch.setenv("RLSIM_NO_EAGER", "1")  # the monolithic launch over all items
        h3 = util.run_gpu(gpu, util.parse(argv), names, seqs, levels)
        monkeypatch.delenv("RLSIM_NO_EAGER")
        assert bool(int(h3.stats().flags) & gpu.F_FUSED) == expect_fused and not (int(h3.stats().flags) & gpu.F_EAGER)
        for k in ("tr", "start", "end", "count0", "count"):
            assert np.array_equal(h.species()[k], h3.species()[k]), k

    # small class, many transcripts per item, one transcript far beyond 4096 molecules, zero-expression gaps
    names, seqs, levels = util.synth_transcriptome(300, seed=41, mean_len=1500, hi=9000, total_molecules=60000, non_acgt=0.0005)
    levels = list(levels)
    levels[5] = 25000
    levels[6] = levels[7] = 0
    both(["-n", "50000", "-c", "4", "-si", "12", "-f", "after_prim_double"], names, seqs, levels)
    both(["-n", "50000", "-c", "4", "-si", "12", "-f", "after_prim"], names, seqs, levels)
    # large class (12-40 kb) + deferred molecules: default -d asks for L/378 breakpoints
    names, seqs, levels = util.synth_transcriptome(24, seed=42, mean_len=22000, sigma=0.4, lo=11000, hi=45000, total_molecules=1500)
    both(["-n", "30000", "-c", "3", "-si", "13", "-f", "after_prim_double"], names, seqs, levels)
    both(["-n", "30000", "-c", "3", "-si", "13", "-f", "after_prim_double:150"], names, seqs, levels)
    # letters outside ACGT in a 20 kb transcript: explicit reverse arrays, still fits the large class
    seqs = list(seqs)
    lens = [len(x) for x in seqs]
    t_mid = int(np.argmin([abs(l - 20000) for l in lens]))
    b = bytearray(seqs[t_mid]); b[5000:5003] = b"NRN"; seqs[t_mid] = bytes(b)
    if lens[t_mid] <= 24000:
        both(["-n", "30000", "-c", "3", "-si", "14", "-f", "after_prim_double"], names, seqs, levels)
    # ... and in a > 24.5 kb one: forward + reverse arrays exceed shared memory -> the library falls back to the HBM path
    t_big = int(np.argmax(lens))
    assert lens[t_big] > 25000
    b = bytearray(seqs[t_big]); b[100] = ord("N"); seqs[t_big] = bytes(b)
    both(["-n", "30000", "-c", "3", "-si", "15", "-f", "after_prim_double"], names, seqs, levels, expect_fused=False, eager=False)


def test_record_layout_opt_in(oracle, gpu, monkeypatch):
    """RLSIM_RECORDS=1: the forward prefix sums of transcripts without letters outside ACGT are kept as one 32-byte record per
    64-entry block (coarse sum, half sum, codes) and the searches recompute weights from the record.  Same prefix sums (probe),
    same fragments as the two-level arrays and as the oracle - dirty transcripts (two-level arrays, second k_intervals launch),
    prim_jump, deferred long molecules and k = 1 included."""
    names, seqs, levels = util.synth_transcriptome(80, seed=51, mean_len=1600, hi=9000, total_molecules=30000, non_acgt=0.0)
    seqs = list(seqs)
    for t in (3, 40):
        b = bytearray(seqs[t]); b[50] = ord("N"); b[200:204] = b"RYKM"; seqs[t] = bytes(b)
    seqs[10] = seqs[10][:6]
    seqs[11] = seqs[11][:70]
    for argv in (["-n", "30000", "-c", "4", "-si", "3", "-f", "after_prim_double"], ["-n", "20000", "-c", "4", "-si", "3", "-f", "after_prim"],
                 ["-n", "5000", "-c", "4", "-si", "3", "-f", "prim_jump:300"], ["-n", "5000", "-c", "3", "-si", "4", "-k", "1"]):
        args = util.parse(argv)
        plain = util.run_gpu(gpu, args, names, seqs, levels)
        monkeypatch.setenv("RLSIM_RECORDS", "1")
        rec = _compare_records(oracle, gpu, argv, names, seqs, levels)
        for k in ("tr", "start", "end", "count0", "count"):
            assert np.array_equal(plain.species()[k], rec.species()[k]), (argv, k)
        pa = int(rec.stats().polya_max)
        for t in (0, 3, 10, 11, 40, 79):
            n = len(seqs[t]) + pa + 1
            for reverse in ((0, 1) if "after_prim_double" in argv else (0,)):
                got = rec.probe_prefix(t, reverse, n)
                monkeypatch.delenv("RLSIM_RECORDS")
                want = plain.probe_prefix(t, reverse, n)
                monkeypatch.setenv("RLSIM_RECORDS", "1")
                assert np.array_equal(got, want), (argv, t, reverse)
        monkeypatch.delenv("RLSIM_RECORDS")
    # deferred long molecules on the record layout
    names, seqs, levels = util.synth_transcriptome(6, seed=8, mean_len=60000, sigma=0.05, lo=50000, hi=70000, total_molecules=600)
    monkeypatch.setenv("RLSIM_RECORDS", "1")
    _compare_records(oracle, gpu, ["-n", "20000", "-c", "5", "-si", "9"], names, seqs, levels)
