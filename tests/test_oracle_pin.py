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


# ---------------------------------------------------------------- the FASTA reader (fasta.go:113-212, input.go:82-124)
def _reader_golden():
    import json
    import os
    return json.load(open(os.path.join(util.GOLDEN, "fasta_reader.json")))


@pytest.mark.parametrize("name", sorted(_reader_golden()))
def test_fasta_reader_pinned_against_release_binary(name):
    """oracle/fasta_reader.py against what the release binary made of the same unusual text: the two report panels the
    reader feeds, the skipped name lines (in order) and the names that reach the FASTA output."""
    from oracle import fasta_reader
    g = _reader_golden()[name]
    mul = float(g["argv"][g["argv"].index("-m") + 1]) if "-m" in g["argv"] else 1.0
    rows = fasta_reader.read(g["text"].encode("latin1"), mul)
    tr_len, expr = fasta_reader.panels(rows)
    assert {str(k): v for k, v in tr_len.items()} == g["panels"]["Transcript lengths"]
    assert {str(k): v for k, v in expr.items()} == g["panels"]["Expression levels"]
    assert [r["line"].decode("latin1") for r in rows if r["name"] is None] == g["skipped"]
    expressed = {r["name"].decode("latin1") for r in rows if r["name"] is not None and r["level"] > 0}
    assert set(g["out_names"]) <= expressed and g["out_names"]


def test_host_mirrors_match_the_reader_oracle():
    """rlsim_b200/fasta.py (vectorised host mirror) against oracle/fasta_reader.py, also on texts the binary cannot take
    (a record without newline makes it panic; the mirrors and the device reader treat it as a bare name line)."""
    from oracle import fasta_reader
    from rlsim_b200 import fasta as hostfasta
    texts = [g["text"].encode("latin1") for g in _reader_golden().values()]
    texts += [b">a$1\nAC\n>>>b$2\nGT\n", b">a$1\nACGT\n>tail$3", b">only", b">a$5\n\n\n", b">a$1\nAC>b$2\nG T\tA\r\n>c", b""]
    for t in texts:
        want = fasta_reader.read(t)
        got = hostfasta.read_records(t)
        assert [(r["line"], r["seq"]) for r in want] == got, t[:40]
        for r in want:
            p = hostfasta.parse_seq_name(r["line"])
            assert (p is None) == (r["name"] is None) and (p is None or (p[0], p[1]) == (r["name"], r["level"])), r["line"]
