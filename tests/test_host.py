"""Host-side logic (no GPU): flags, mixture language, FASTA reader, report schema, C-ABI surface, sharding."""
import ctypes
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import util
from rlsim_b200 import args as A
from rlsim_b200 import fasta, mixture, report, shard

ROOT = os.path.dirname(util.HERE)


def test_default_flags_match_reference():
    a = util.parse(["-n", "10"])
    assert (a.nr_cycles, a.strand_bias, a.priming_temp, a.kmer_length, a.fixed_eff) == (11, 0.5, 5.0, 6, 0.0)
    assert a.frag_method.name == "after_prim_double" and a.frag_method.param == 0
    assert len(a.raw_gc_effs) == 101 and a.raw_gc_effs[17] == 0.6214855142919264  # args.go:194
    assert str(a.target_mix[0]) == "sn:(189, 24, -1.09975, 76, 294)"
    assert str(a.polya_param[0]) == "g:(150, 1, 0, 220)"
    assert a.len_eff_param.shape == 0.0 and a.gob_dir.startswith("rlsim_gob_")
    p = a.to_params()
    assert p.gc_mode == 1 and p.n_target == 1 and p.n_polya == 1 and p.kmer_len == 6


def test_go_flag_syntax():
    a = A.CmdArgs.parse(["--n=5", "-c", "3", "-g", "-v", "-f=pre_prim", "x.fa", "-n", "7"])
    assert a.req_frags == 5 and a.nr_cycles == 3 and a.verbose and a.gob_dir == ""
    assert a.frag_method.param == 2000  # args.go:301-304
    assert a.input_files == ["x.fa", "-n", "7"]  # parsing stops at the first non-flag argument
    with pytest.raises(A.ArgError):
        A.CmdArgs.parse(["-nope", "1"])
    with pytest.raises(A.ExitRequest):
        A.CmdArgs.parse(["-V"])
    with pytest.raises(A.ExitRequest):
        A.CmdArgs.parse(["x.fa"])  # no fragments requested


@pytest.mark.parametrize("bad", ["1.0:sn:(189,24,76,294)", "1.0:x:(1,2,3,4)", "1.0:n:(1,2,9,3)", "", "1.0:n:1,2,3,4",
                                 "1.0:n:(-1,2,3,4)"])
def test_mixture_errors(bad):
    with pytest.raises(mixture.ArgError):
        mixture.parse_mixture(bad)


def test_mixture_parse():
    m = mixture.parse_mixture("0.9:sn:(600,50,4,300,2000) + 0.1:n:(700, 1, 600, 2000)")
    assert [c.type for c in m] == ["sn", "n"] and m[0].shape == 4 and m[1].weight == 0.1
    assert mixture.global_min_max(m) == (300, 2000)


def test_eff_param_and_method_errors():
    assert A.parse_eff_param(" (1.0,0.5,0.8) ").min == 0.5
    for bad in ["1,2,3", "(1,2)", "(-1,0.5,0.8)", "(1,1.5,0.8)"]:
        with pytest.raises(A.ArgError):
            A.parse_eff_param(bad)
    with pytest.raises(A.ArgError):
        A.parse_frag_method("bogus")
    assert A.parse_frag_method("prim_jump:300").param == 300


def test_raw_params_file():
    a = util.parse(["-j", util.RAW_PARAMS])
    assert a.req_frags == 964058 and a.nr_cycles == 12 and len(a.raw_gc_effs) == 101
    assert abs(sum(a.raw_len_probs.probs) - 1.0) < 1e-12
    p = a.to_params()
    assert p.n_target == 0 and p.gc_mode == 1
    a2 = util.parse(["-j", util.RAW_PARAMS, "-jm", "0.5", "-n", "5"])
    assert min(a2.processed_raw_gc_effs()) == 0.5 and a2.req_frags == 5


def test_fasta_reader_quirks():
    data = b">a$5 \nacgt\nNN tt\r\n>b$x\nAC\n>c\nAC\n>d$7\nA>C\n"
    recs = fasta.read_records(data)
    assert recs[0] == (b"a$5 ", b"ACGTNNTT")
    # fasta.go:138: any '>' opens a record once the buffer holds more than one byte
    assert [r[0] for r in recs] == [b"a$5 ", b"b$x", b"c", b"d$7", b"C"]
    assert fasta.parse_seq_name(b"a$5 ") == (b"a", 5)
    assert fasta.parse_seq_name(b"b$x") is None and fasta.parse_seq_name(b"c") is None and fasta.parse_seq_name(b"$3") is None
    # input.go:92-93: TrimSpace + Sscanf("%d") into a uint64 - no sign, must fit 64 bits, trailing text ignored
    assert fasta.parse_seq_name(b"a$+5") is None and fasta.parse_seq_name(b"a$-5") is None
    assert fasta.parse_seq_name(b"a$\t\v\f 12junk\r") == (b"a", 12)
    assert fasta.parse_seq_name(b"a$18446744073709551615") == (b"a", 2**64 - 1) and fasta.parse_seq_name(b"a$18446744073709551616") is None
    assert fasta.parse_seq_name(b"a$1$2") is None
    names, seqs, levels = fasta.load_transcripts([util.BASIC_FASTA], 1.0)
    assert [len(s) for s in seqs] == [1075, 3477, 4002, 2729, 1690] and levels == [10200, 8900, 2300, 2160, 0]
    assert fasta.load_transcripts([util.BASIC_FASTA], 0.1)[2] == [1020, 890, 230, 216, 0]  # input.go:113


def test_report_schema():
    with tempfile.TemporaryDirectory() as td:
        r = report.Report(os.path.join(td, "r.json"))
        r.report_map({100: 3, 7: 1}, "Length", "Count", "Target lenghts", "bar")
        r.report_float([0.0, 0.5], [0.1, 0.2], "GC content (%)", "Efficiency", "GC efficiency function", "line")
        r.write_json()
        d = json.load(open(r.path))
    assert d["Target lenghts"] == {"Xl": "Length", "Yl": "Count", "Vis": "bar", "Pos": 0, "Data": {"100": 3, "7": 1}}
    assert d["GC efficiency function"]["Pos"] == 1 and d["GC efficiency function"]["Data"] == {"0": 0.1, "0.5": 0.2}


def test_c_abi_exports_every_declared_symbol():
    """The shared library loads without a GPU and exports exactly what include/rlsim_gpu.h declares."""
    from rlsim_b200 import _native
    lib = _native.load()
    hdr = open(os.path.join(ROOT, "include", "rlsim_gpu.h")).read()
    declared = sorted(set(re.findall(r"\b(rlsim_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 25
    for name in declared:
        getattr(lib, name)
    assert sorted(declared) == sorted(_native.EXPORTS)


def test_struct_layout_matches_header():
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "rlsim_gpu.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu",' \
          "sizeof(rlsim_params),sizeof(rlsim_mixcomp),sizeof(rlsim_stats),offsetof(rlsim_params,gc_raw)," \
          "offsetof(rlsim_params,target),offsetof(rlsim_params,seed_init));return 0;}"
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    from rlsim_b200._native import Stats
    assert sizes == [ctypes.sizeof(A.CParams), ctypes.sizeof(A.CMixComp), ctypes.sizeof(Stats), A.CParams.gc_raw.offset,
                     A.CParams.target.offset, A.CParams.seed_init.offset]
    from oracle import pyoracle
    assert ctypes.sizeof(pyoracle.Params) == ctypes.sizeof(A.CParams)


def test_no_cpu_fallback_without_gpu():
    """On a box without a CUDA device the product fails loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from rlsim_b200 import _native
    with pytest.raises(_native.RlsimError):
        _native.Handle(0)


def test_product_never_touches_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rlsim_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "oracle/" not in txt.replace("the oracle", ""), f


def _go_methods(src, receiver):
    """{method: (params, returns)} of the methods go/gpu_pool.go declares on `receiver` (value or pointer)."""
    out = {}
    for m in re.finditer(r"^func \(\w+ \*?%s\) (\w+)\((.*?)\)\s*(.*?)\s*\{" % receiver, src, re.M):
        out[m.group(1)] = (m.group(2).strip(), m.group(3).strip())
    return out


def test_go_shim_implements_the_reference_interfaces():
    """go/gpu_pool.go: GpuPool has all 11 Pooler methods (pool.go:36-48) and GpuSampler both Sampler methods
    (sampler.go:39-42), with the reference's parameter lists and result types.  No Go toolchain here: the file is also
    checked for balanced delimiters, that every C symbol it calls is declared in include/rlsim_gpu.h, and that the
    committed method lists are still what the reference declares (when /root/reference is mounted)."""
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "go_interfaces.json")))
    assert len(want["Pooler"]) == 11 and len(want["Sampler"]) == 2
    if os.path.isdir("/root/reference/src"):
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        import make_go_interfaces
        assert make_go_interfaces.extract() == want
    src = open(os.path.join(ROOT, "go", "gpu_pool.go")).read()
    norm = lambda x: re.sub(r"\s+", " ", x.strip())
    for iface, typ in (("Pooler", "GpuPool"), ("Sampler", "GpuSampler")):
        have = _go_methods(src, typ)
        for name, sig in want[iface].items():
            assert name in have, "%s.%s is missing" % (typ, name)
            types = lambda plist: [norm(x).split(" ", 1)[1] for x in plist.split(",") if x.strip()]  # parameter names are free
            assert types(have[name][0]) == types(sig["params"]), (name, have[name][0], sig["params"])
            assert norm(have[name][1]) == norm(sig["returns"]), (name, have[name][1], sig["returns"])
    assert "var _ Pooler = (*GpuPool)(nil)" in src and "var _ Sampler = GpuSampler{}" in src
    code = re.sub(r"//[^\n]*", "", src)                      # comments ...
    code = re.sub(r'"(\\.|[^"\\])*"', '""', code)            # ... and string literals out of the way
    for a, b in ("()", "{}", "[]"):
        assert code.count(a) == code.count(b), "unbalanced %s%s" % (a, b)
    hdr = open(os.path.join(ROOT, "include", "rlsim_gpu.h")).read()
    for sym in set(re.findall(r"C\.(rlsim_\w+|RLSIM_\w+)", src)):
        assert re.search(r"\b%s\b" % sym, hdr), "C.%s is not declared in include/rlsim_gpu.h" % sym
    # the report bookkeeping the reference does in SampleFragments / SampleTranscripts (sampler.go:95-101,149-153)
    for call in ("st.UpdateSampled(", "st.UpdateMissing(", "st.UpdateAfterSampling(", "st.LogSamplingRatio(", "st.UpdateNrSampled("):
        assert call in src, call


def test_go_spec_files_restate_the_oracle():
    """go/philox.go and go/detmath.go carry the same constants as the plain-C text of the spec (oracle/philox.h,
    oracle/detmath.h): Philox multipliers / Weyl constants / key-derivation words / tags, and every floating-point literal
    of the det_* functions.  (They cannot be compiled here; this guards against transcription slips.)"""
    def floats(text):
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r"//[^\n]*", "", text)
        out = set()
        for tok in re.findall(r"(?<![\w.])-?\d+\.\d*(?:[eE][-+]?\d+)?|(?<![\w.])-?\d+[eE][-+]?\d+", text):
            out.add(float(tok))
        return out
    go_px = open(os.path.join(ROOT, "go", "philox.go")).read()
    c_px = open(os.path.join(ROOT, "oracle", "philox.h")).read()
    for const in ("0xD2511F53", "0xCD9E8D57", "0x9E3779B9", "0xBB67AE85", "0x524c5349", "0x4d423230"):
        assert const.lower() in go_px.lower() and const.lower() in c_px.lower(), const
    for tag, val in (("PxTagTarget", 1), ("PxTagFrag", 2), ("PxTagPcr", 3), ("PxTagSelect", 4), ("PxTagStrand", 5)):
        assert re.search(r"%s\s*=\s*%d\b" % (tag, val), go_px), tag
    assert "1.1102230246251565404e-16" in go_px and "2.2204460492503130808e-16" in go_px
    assert "bits.Mul64" in go_px
    go_dm = floats(open(os.path.join(ROOT, "go", "detmath.go")).read())
    c_dm = floats(open(os.path.join(ROOT, "oracle", "detmath.h")).read())
    missing = sorted(x for x in c_dm if x not in go_dm and -x not in go_dm)
    assert not missing, missing
    for src in (go_px, open(os.path.join(ROOT, "go", "detmath.go")).read()):
        code = re.sub(r"//[^\n]*", "", src)
        for a, b in ("()", "{}", "[]"):
            assert code.count(a) == code.count(b)


def test_exchange_plan_host_arithmetic():
    """rlsim_exchange_plan: where each rank's groups start in its send buffer and where each sender's records land
    (the routing arithmetic of rlsim_sample_distributed; no GPU involved)."""
    from rlsim_b200 import _native
    rng = np.random.default_rng(4)
    for n in (1, 2, 3, 8):
        c = rng.integers(0, 1000, size=(n, n)).astype(np.uint64)
        c[0, n - 1] = 0
        for r in range(n):
            so, ro = _native.exchange_plan(c, r)
            assert list(so) == [int(c[r, :d].sum()) for d in range(n + 1)]
            assert list(ro) == [int(c[:s, r].sum()) for s in range(n + 1)]
        # what every rank sends to rank r is what rank r expects to receive
        assert all(int(_native.exchange_plan(c, r)[1][-1]) == int(c[:, r].sum()) for r in range(n))
    with pytest.raises(_native.RlsimError):
        _native.exchange_plan(np.zeros((2, 2), np.uint64), 2)


def test_contiguous_partition():
    rng = np.random.default_rng(0)
    w = rng.gamma(0.5, 10.0, 1000)
    for n in (1, 2, 3, 8):
        b = shard.contiguous_partition(w, n)
        assert b[0] == 0 and b[-1] == 1000 and all(x <= y for x, y in zip(b, b[1:])) and len(b) == n + 1
        loads = [w[b[i]:b[i + 1]].sum() for i in range(n)]
        assert max(loads) <= w.sum() / n + w.max() + 1e-9
    assert shard.contiguous_partition([0, 0, 0], 2) == [0, 0, 3]


_GLOO = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch.distributed as dist
from rlsim_b200 import shard
dist.init_process_group("gloo")
r = dist.get_rank()
local = np.array([10 * (r + 1), 0, 7, 2**40 + r], dtype=np.uint64)
tot, base = shard.allgather_totals(local)
assert list(tot) == [30, 0, 14, 2**41 + 1], tot
assert list(base) == ([0, 0, 0, 0] if r == 0 else [10, 0, 7, 2**40]), base
h = shard.allreduce_hist(np.array([r + 1, 5], dtype=np.uint64))
assert list(h) == [3, 10]
big = np.array([2**63 + 5], dtype=np.uint64)
try:
    shard.allgather_totals(big)
    raise SystemExit("overflow not detected")
except OverflowError:
    pass

# the selection exchange (RecordExchange): [totals | partial target] in one all-gather, then counts + records all-to-all
import ctypes
class FakeHandle:
    def class_totals(self):
        return np.array([5 + r, 3, 0, 7], dtype=np.uint64)
    def hist(self, kind, nb):
        assert kind == 0
        return (np.arange(nb) + r).astype(np.uint64)
    def select_classes(self, all_tot, rank, ptr, cap):
        assert all_tot.shape == (2, 4) and list(all_tot[:, 0]) == [5, 6] and rank == r
        counts = np.array([r + 1, 2], dtype=np.uint64)      # records for rank 0, rank 1
        needed = int(counts.sum())
        if needed <= cap:
            rec = (ctypes.c_int64 * (2 * needed)).from_address(ptr)
            k = 0
            for dest in range(2):
                for j in range(int(counts[dest])):
                    rec[2 * k], rec[2 * k + 1] = 100 * r + dest, j
                    k += 1
        return counts, needed
    def apply_selected(self, ptr, n):
        rec = (ctypes.c_int64 * (2 * n)).from_address(ptr) if n else []
        self.got = sorted((rec[2 * i], rec[2 * i + 1]) for i in range(n))
    def finish_sample(self):
        self.finished = True
class FakePool:
    _target_nb = 3
    def _sum_target(self, tot):
        self.target, self._target_nb = list(tot), 0
fh, fp = FakeHandle(), FakePool()
all_tot = shard.RecordExchange(None).run(fh, r, 2, 16, pool=fp)
assert all_tot.shape == (2, 4) and list(all_tot[1]) == [6, 3, 0, 7]
assert fp.target == [1, 3, 5] and fp._target_nb == 0                       # (0,1,2) + (1,2,3)
want = {0: [(0, 0), (100, 0), (100, 1)], 1: [(1, 0), (1, 1), (101, 0), (101, 1)]}[r]   # from rank 0: r+1=1 / 2; from rank 1: 2 / 2
assert fh.got == sorted(want) and fh.finished, (r, fh.got)
dist.destroy_process_group()
"""


def test_allgather_world_size_2_gloo():
    """The N>1 host path: per-length totals all-gathered, rank bases derived, overflow detected (fragstats.go:88-94)."""
    with tempfile.TemporaryDirectory() as td:
        script = os.path.join(td, "w.py")
        open(script, "w").write(_GLOO % ROOT)
        env = dict(os.environ, MASTER_ADDR="127.0.0.1")
        subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                               "--master-addr", "127.0.0.1", "--master-port", "29731", script], env=env, timeout=300)


def test_cpp_twin_cli_surface():
    """rlsim_gpu: -V, fatal argument errors (exit 1, reference messages) and loud failure without a GPU."""
    import torch
    exe = os.path.join(ROOT, "rlsim_b200", "rlsim_gpu")
    if not os.path.exists(exe):
        from rlsim_b200 import build
        build.build()
    assert subprocess.check_output([exe, "-V"]).decode().strip() == "1.3"
    for argv, msg in ((["-n", "10", "-f", "bogus"], b"Malformed fragmentation method string"),
                      (["-n", "10", "-d", "1:n:(1,2,9,3)"], b"Maximum fragment size is smaller"),
                      (["-n", "10", "-flg", "2"], b"fragment loss probability"),
                      (["-n", "10", "-eg", "(1,2,3)"], b"outside the range"),
                      (["-nope", "1"], b"flag provided but not defined")):
        p = subprocess.run([exe] + argv + [util.BASIC_FASTA], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert p.returncode == 1 and msg in p.stderr, (argv, p.stderr)
    if not torch.cuda.is_available():
        p = subprocess.run([exe, "-n", "10", util.BASIC_FASTA], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert p.returncode == 1 and b"cuda" in p.stderr.lower()


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's own CPU implementation on a bounded sample) prints ONE JSON line with
    the keys the driver reads; under torchrun only rank 0 prints."""
    env = dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--scale", "0.02"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=600)
    assert p.returncode == 0, p.stderr.decode()[-1500:]
    lines = [l for l in p.stdout.decode().splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "fragments/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"] == "fragments/sec (frag+PCR+size-select)" and d["vs_baseline"] is None
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"] == {"value": d["value"], "unit": "fragments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    env["RANK"] = "1"
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--scale", "0.02"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=600)
    assert p.returncode == 0 and p.stdout.strip() == b""


def test_bench_own_arm_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--scale", "0.01"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert p.returncode != 0 and b"no CPU fallback" in p.stderr and p.stdout.strip() == b""


def test_pack_bases_layout():
    """The packed host format of rlsim_add_transcripts_packed (include/rlsim_gpu.h): 2 bits per base, low bits first; the
    mask marks every letter outside ACGT and is None for a clean sequence."""
    from rlsim_b200 import _native
    rng = np.random.default_rng(3)
    b = rng.choice(np.frombuffer(b"ACGTNRY", dtype=np.uint8), size=1003)
    codes, nmask = _native.pack_bases(b)
    assert codes.dtype == np.uint8 and len(codes) == 251 and len(nmask) == 126
    i = np.arange(len(b))
    c = (codes[i >> 2] >> (2 * (i & 3))) & 3
    m = (nmask[i >> 3] >> (i & 7)) & 1
    clean = np.isin(b, np.frombuffer(b"ACGT", dtype=np.uint8))
    assert np.array_equal(m.astype(bool), ~clean)
    assert np.array_equal(np.frombuffer(b"ACGT", dtype=np.uint8)[c][clean], b[clean]) and not c[~clean].any()
    assert _native.pack_bases(np.frombuffer(b"ACGTTGCA", dtype=np.uint8))[1] is None
    assert _native.pack_bases(np.frombuffer(b"CA", dtype=np.uint8))[0].tolist() == [1]


def test_prefix_kernel_uses_bulk_async_copies():
    """The fine prefix sums leave shared memory through the TMA engine (cp.async.bulk.global.shared::cta): the sm_100a SASS of
    the built library carries UBLKCP (profiling guide: the mnemonic that proves bulk asynchronous copies)."""
    import shutil
    import subprocess
    from rlsim_b200 import build
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([tool, "-sass", "-fun", "_ZN2rl8k_prefixILb1EEEvNS_14DevTranscriptsEjjidPKjS3_PKyPjS6_PmS7_PNS_5Rec32EiS6_S6_j", build.LIB],
                          capture_output=True, text=True).stdout
    if "UBLKCP" not in sass:  # symbol name drift: look at the whole library
        sass = subprocess.run([tool, "-sass", build.LIB], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass


def test_streaming_host_copy_is_a_memcpy(tmp_path):
    """csrc/hostcopy.cpp (staging of pageable caller buffers with non-temporal stores): every byte arrives, nothing beside
    the destination is touched, for all alignments of source and destination and sizes around the vector width."""
    import shutil
    import subprocess
    if not shutil.which("g++"):
        pytest.skip("g++ not available")
    so = str(tmp_path / "hostcopy.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", os.path.join(ROOT, "rlsim_b200", "csrc", "hostcopy.cpp"), "-o", so])
    lib = ctypes.CDLL(so)
    lib.rl_stream_copy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    rng = np.random.default_rng(1)
    for n in (0, 1, 31, 127, 128, 4095, 4096, 4097, 4096 + 127, 100003, 1_000_001):
        for off_s in (0, 1, 7, 31):
            for off_d in (0, 3, 17, 32):
                src = rng.integers(0, 256, n + 96, dtype=np.uint8)
                dst = np.zeros(n + 96, np.uint8)
                lib.rl_stream_copy(dst.ctypes.data + off_d, src.ctypes.data + off_s, n)
                assert np.array_equal(dst[off_d:off_d + n], src[off_s:off_s + n]), (n, off_s, off_d)
                assert not dst[:off_d].any() and not dst[off_d + n:].any(), (n, off_s, off_d)
