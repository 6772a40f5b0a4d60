"""The host twin end to end on the GPU: `python -m rlsim_b200 <rlsim flags> transcripts.fas` -> FASTA + JSON report."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(util.HERE)


def _run_cli(argv, tmp_path):
    rep = str(tmp_path / "report.json")
    p = subprocess.run([sys.executable, "-m", "rlsim_b200", "-r", rep] + argv + [util.BASIC_FASTA], cwd=ROOT,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    return p.stdout, json.load(open(rep)), p.stderr.decode()


def test_cli_matches_oracle_and_report_schema(oracle, tmp_path):
    argv = ["-n", "10000", "-si", "42", "-v"]
    fasta, rep, log = _run_cli(argv, tmp_path)
    args = util.parse(argv)
    names, seqs, levels = util.transcripts(args)
    o = util.run_oracle(oracle, args, names, seqs, levels, oracle.MODE_PHILOX)
    assert fasta == o.fasta(names)  # FASTA text identical to the oracle's records
    # the twelve panels of the reference report, in the reference's order
    titles = sorted(rep, key=lambda k: rep[k]["Pos"])
    assert titles == ["Target lenghts", "GC efficiency function", "Length efficiency function", "Fragdist after fragmentation",
                      "Fragdist after PCR", "Fragdist after sampling", "Missing fragments", "Sampled fragments",
                      "Poly-A tail lengths", "Transcript lengths", "Expression levels", "Sampling ratio"]
    for t in titles:
        assert set(rep[t]) == {"Xl", "Yl", "Vis", "Pos", "Data"}
    pi = util.panel_int
    assert pi(rep["Target lenghts"]["Data"]) == o.hist_dict(oracle.H_TARGET)
    assert pi(rep["Fragdist after fragmentation"]["Data"]) == o.hist_dict(oracle.H_AFTER_FRAG)
    assert pi(rep["Fragdist after PCR"]["Data"]) == o.hist_dict(oracle.H_AFTER_PCR)
    assert pi(rep["Poly-A tail lengths"]["Data"]) == o.hist_dict(oracle.H_POLYA)
    assert {k: v for k, v in pi(rep["Sampled fragments"]["Data"]).items() if v} == o.hist_dict(oracle.H_SAMPLED)
    assert sum(pi(rep["Missing fragments"]["Data"]).values()) == 0
    # deterministic panels equal the release binary's (golden fixture)
    g = util.golden("c1")["panels"]
    assert rep["Transcript lengths"]["Data"] == g["Transcript lengths"]
    assert rep["Expression levels"]["Data"] == g["Expression levels"]
    assert rep["GC efficiency function"]["Data"] == g["GC efficiency function"]
    assert rep["Length efficiency function"]["Data"] == g["Length efficiency function"]
    assert "Sampled 10000 fragments." in log and "Missing fragments: 0" in log


def test_cli_fixed_efficiency_omits_efficiency_panels(tmp_path):
    _, rep, _ = _run_cli(["-n", "500", "-e", "0.7", "-si", "3", "-m", "0.1"], tmp_path)
    assert len(rep) == 10 and "GC efficiency function" not in rep  # thermocycler.go:196-198


def test_cli_errors_are_fatal(tmp_path):
    p = subprocess.run([sys.executable, "-m", "rlsim_b200", "-n", "10", "-f", "bogus", util.BASIC_FASTA], cwd=ROOT,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 1 and b"Malformed fragmentation method string" in p.stderr
    p = subprocess.run([sys.executable, "-m", "rlsim_b200", "-n", "10", "-k", "20", util.BASIC_FASTA], cwd=ROOT,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 1 and b"primer length" in p.stderr


def test_pcr_stress_large_counts(oracle, gpu):
    """Config C5 in miniature: 30 cycles at high efficiency -> copy numbers ~1e8 per species, BTRS with huge n,
    u64 bookkeeping (pool total ~1e12); still bit for bit."""
    args = util.parse(["-n", "50000", "-c", "30", "-eg", "(1.0,0.9,1.0)", "-m", "0.05", "-si", "11"])
    names, seqs, levels = util.transcripts(args)
    o = util.run_oracle(oracle, args, names, seqs, levels, oracle.MODE_PHILOX, stages="pcr")
    h = util.run_gpu(gpu, args, names, seqs, levels, stages="pcr")
    so, sg = o.species(), h.species()
    assert np.array_equal(so["count"], sg["count"]) and so["count"].max() > 10**8
    assert h.stats().pool_total == o.pool_total > 10**11


def test_pcr_overflow_is_reported(gpu):
    """thermocycler.go:142-144: u64 wrap of a copy number is fatal."""
    args = util.parse(["-n", "10", "-c", "70", "-e", "1.0", "-m", "0.001", "-si", "1"])
    names, seqs, levels = util.transcripts(args)
    with pytest.raises(gpu.RlsimError) as e:
        util.run_gpu(gpu, args, names, seqs, levels, stages="pcr")
    assert "Integer overflow detected when amplifying fragments!" in str(e.value)


def _run_cpp(argv, tmp_path):
    exe = os.path.join(ROOT, "rlsim_b200", "rlsim_gpu")
    rep = str(tmp_path / "report_cpp.json")
    p = subprocess.run([exe, "-r", rep] + argv + [util.BASIC_FASTA], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    return p.stdout, json.load(open(rep)), p.stderr.decode()


@pytest.mark.parametrize("argv", [["-n", "10000", "-si", "42", "-v"], ["-j", util.RAW_PARAMS, "-n", "3000", "-si", "5", "-f", "after_prim"],
                                  ["-n", "2000", "-si", "9", "-e", "0.6", "-b", "0.2", "-m", "0.3", "-f", "pre_prim:800"]])
def test_cpp_twin_equals_python_twin(argv, tmp_path):
    """rlsim_gpu (C++ host over the C ABI) and `python -m rlsim_b200` are the same program: identical FASTA, identical
    report (numbers compared as numbers)."""
    f_cpp, r_cpp, log = _run_cpp(argv, tmp_path)
    f_py, r_py, _ = _run_cli(argv, tmp_path)
    assert f_cpp == f_py
    assert sorted(r_cpp) == sorted(r_py)
    for t in r_py:
        assert {k: r_cpp[t][k] for k in ("Xl", "Yl", "Vis", "Pos")} == {k: r_py[t][k] for k in ("Xl", "Yl", "Vis", "Pos")}, t
        a, b = r_cpp[t]["Data"], r_py[t]["Data"]
        assert a.keys() == b.keys(), t
        for k in a:
            assert a[k] == pytest.approx(b[k], rel=1e-12), (t, k)
    if "-v" in argv:
        assert "Sampled 10000 fragments." in log


def test_cpp_twin_fractional_length_efficiency(tmp_path):
    """With -el the C++ host hands its own length-efficiency table to the library (rlsim_set_len_eff), the way a Go
    host would: the run completes and the table in the report is the one used."""
    f_cpp, r_cpp, _ = _run_cpp(util.C2[:2] + ["-si", "4"] + util.C2[2:], tmp_path)
    assert len(r_cpp) == 12 and len(r_cpp["Length efficiency function"]["Data"]) == 1701
    g = util.golden("c2")["panels"]["Length efficiency function"]
    for k, v in g.items():
        assert r_cpp["Length efficiency function"]["Data"][k] == pytest.approx(v, rel=1e-14)
    assert f_cpp.count(b">") == sum(r_cpp["Sampled fragments"]["Data"].values())
