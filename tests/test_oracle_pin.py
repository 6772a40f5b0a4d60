"""Pin the CPU oracle (ref mode) against the reference's own release binary.

The committed fixtures tests/golden/*.json were produced by tests/golden/make_golden.py from
releases/rlsim-1.4_amd64.tar.gz:bin/rlsim.  With a fixed -si the binary reproduces its target-length,
poly-A and after-fragmentation histograms run to run; the oracle - Go math/rand replica + restated
nnthermo / fragmentor / random / target code - must reproduce them EXACTLY (every bin of every histogram).
"""
import math

import pytest

import util


def _oracle_panels(oracle, name):
    g = util.golden(name)
    args = util.parse(g["argv"] + ["-si", str(g["seed"])])
    names, seqs, levels = util.transcripts(args)
    s = util.run_oracle(oracle, args, names, seqs, levels, oracle.MODE_REF, stages="fragment")
    return g, args, s, (names, seqs, levels)


@pytest.mark.parametrize("name", sorted(util.PIN_CONFIGS))
def test_ref_mode_reproduces_release_binary(oracle, name):
    g, args, s, _ = _oracle_panels(oracle, name)
    assert s.hist_dict(oracle.H_TARGET) == util.panel_int(g["panels"]["Target lenghts"])
    assert s.hist_dict(oracle.H_POLYA) == util.panel_int(g["panels"]["Poly-A tail lengths"])
    assert s.hist_dict(oracle.H_AFTER_FRAG) == util.panel_int(g["panels"]["Fragdist after fragmentation"])


@pytest.mark.parametrize("name", sorted(util.PIN_CONFIGS))
def test_input_statistics(name):
    g = util.golden(name)
    args = util.parse(g["argv"])
    names, seqs, levels = util.transcripts(args)
    trl, exl = {}, {}
    for s, lv in zip(seqs, levels):
        trl[len(s)] = trl.get(len(s), 0) + lv
        exl[lv] = exl.get(lv, 0) + 1
    assert trl == util.panel_int(g["panels"]["Transcript lengths"])
    assert exl == util.panel_int(g["panels"]["Expression levels"])


@pytest.mark.parametrize("name", sorted(util.PIN_CONFIGS))
def test_efficiency_functions(oracle, name):
    """GC panel: bit-exact for raw tables and integer shapes; <= 4 ulp otherwise (Go's asm Exp/Log are not
    reproducible).  Length panel: <= 4 ulp (math.Pow with a fractional exponent)."""
    g = util.golden(name)
    args = util.parse(g["argv"])
    p = oracle.Params.from_buffer_copy(bytes(args.to_params()))
    L = oracle.lib()
    import ctypes
    gcp = {float(k): v for k, v in g["panels"]["GC efficiency function"].items()}
    gc = 0.0
    exact = p.gc_mode == 1 or float(p.gc_shape).is_integer()
    for _ in range(201):
        for det in (0, 1):
            got = L.orc_gc_eff_fixed(ctypes.byref(p), gc / 100.0, det)
            want = gcp[float("%g" % gc)]
            if exact:
                assert got == want, (gc, got, want)
            else:
                assert math.isclose(got, want, rel_tol=1e-15, abs_tol=0), (gc, got, want)
        gc += 0.5
    lp = util.panel_int(g["panels"]["Length efficiency function"])
    low, high = min(lp), max(lp)
    import numpy as np
    for det in (0, 1):
        out = np.zeros(high - low + 1)
        L.orc_len_eff_table(ctypes.byref(p), low, high, det, out.ctypes.data)
        for l in range(low, high + 1):
            assert math.isclose(out[l - low], lp[l], rel_tol=1e-15, abs_tol=0), (l, out[l - low], lp[l])


def test_live_binary_matches_golden(oracle):
    """When the release binary is present (build container, GPU box) re-run it and compare with the fixture."""
    from oracle import refbin
    if not refbin.available():
        pytest.skip("oracle/_ref release binary not unpacked")
    g = util.golden("c1")
    r = refbin.run(g["argv"] + ["-si", str(g["seed"])], util.BASIC_FASTA, keep_fasta=True)
    for p in ("Target lenghts", "Poly-A tail lengths", "Fragdist after fragmentation"):
        assert r["report"][p]["Data"] == g["panels"][p]
    # FASTA contract (frag.go:71-126): every record is plus[start:end] or its reverse complement
    args = util.parse(g["argv"])
    names, seqs, levels = util.transcripts(args)
    byname = {n: s + b"A" * 220 for n, s in zip(names, seqs)}
    recs = refbin.parse_fasta(r["fasta"])
    assert len(recs) == 10000
    import numpy as np
    L = oracle.lib()
    for fid, name, strand, start, end, seq in recs[:2000]:
        ref = byname[name][start:end]
        if strand == "-":
            out = np.zeros(len(ref), dtype=np.uint8)
            L.orc_revcomp(ref, len(ref), out.ctypes.data)
            ref = out.tobytes()
        assert seq == ref
