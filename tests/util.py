"""Shared helpers of the test-suite: configurations, fixtures, synthetic transcriptomes."""
import json
import os

import numpy as np

from rlsim_b200.args import CmdArgs
from rlsim_b200.fasta import load_transcripts

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
BASIC_FASTA = os.path.join(GOLDEN, "test_transcripts.fas")  # src/test/basic/test_transcripts.fas of the reference
RAW_PARAMS = os.path.join(GOLDEN, "raw_params.json")        # src/test/basic/raw_params.json of the reference

C1 = ["-n", "10000"]
C2 = ["-n", "500000", "-d", "0.9:sn:(600,50,4,300,2000) + 0.1:n:(700,1,600,2000)", "-f", "after_prim_double", "-c", "15",
      "-eg", "(1.0,0.5,0.8)", "-el", "(0.1,0.7,1.0)"]

# name -> argv (without seed / input); every entry is pinned against the release binary in tests/golden/
PIN_CONFIGS = {
    "c1": C1,
    "c2": C2,
    "after_prim": ["-n", "5000", "-f", "after_prim"],
    "after_noprim": ["-n", "5000", "-f", "after_noprim"],
    "after_noprim_double": ["-n", "5000", "-f", "after_noprim_double"],
    "pre_prim": ["-n", "5000", "-f", "pre_prim"],
    "pre_prim_1500": ["-n", "5000", "-f", "pre_prim:1500"],
    "prim_jump": ["-n", "5000", "-f", "prim_jump"],
    "prim_jump_300": ["-n", "5000", "-f", "prim_jump:300"],
    "after_prim_double_250": ["-n", "5000", "-f", "after_prim_double:250"],
    "loss_k8": ["-n", "5000", "-flg", "0.3", "-a", "1.0:n:(100,30,0,200)", "-k", "8", "-p", "2.0"],
    "polya_gamma_mix": ["-n", "5000", "-a", "0.5:g:(80,4,0,250) + 0.5:g:(20,0.5,0,250)", "-m", "0.1"],
    "basic_makefile": ["-d", "1:n:(300,20,100,5000)", "-eg", "(8.0,0.5,0.95)", "-el", "(0.4,0.5,1.0)", "-n", "20000"],
}
PIN_SEEDS = {"c1": 42, "c2": 42}
PIN_PANELS = ["Target lenghts", "Poly-A tail lengths", "Fragdist after fragmentation", "Transcript lengths",
              "Expression levels", "GC efficiency function", "Length efficiency function"]


def pin_seed(name):
    return PIN_SEEDS.get(name, 7)


def parse(argv, fasta=BASIC_FASTA):
    return CmdArgs.parse(list(argv) + [fasta])


def transcripts(args):
    return load_transcripts(args.input_files, args.expr_mul)


def golden(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)


def panel_int(d):
    return {int(k): v for k, v in d.items()}


def synth_transcriptome(n, seed=20261019, mean_len=1700.0, sigma=0.75, lo=200, hi=30000, total_molecules=None,
                        non_acgt=0.0):
    """SURVEY 8(d) C3 generator: lognormal lengths, per-transcript GC ~ Beta(8,8) in [0.3,0.7], iid bases,
    expression from the `sel` default gamma mixture 0.5:g:(5e3,0.5) + 0.5:g:(5e4,1000), rescaled."""
    rng = np.random.default_rng(seed)
    lens = np.clip(np.floor(rng.lognormal(np.log(mean_len), sigma, n)), lo, hi).astype(np.int64)
    gc = 0.30 + 0.40 * rng.beta(8, 8, n)
    comp = rng.random(n) < 0.5
    expr = np.where(comp, rng.gamma(0.5, 5e3 / 0.5, n), rng.gamma(1000.0, 5e4 / 1000.0, n))
    if total_molecules is not None:
        expr = expr * (total_molecules / expr.sum())
    levels = np.floor(expr).astype(np.uint64)
    seqs = []
    for i in range(n):
        p = np.array([(1 - gc[i]) / 2, gc[i] / 2, gc[i] / 2, (1 - gc[i]) / 2])
        codes = rng.choice(4, size=lens[i], p=p)
        s = np.frombuffer(b"ACGT", dtype=np.uint8)[codes].copy()
        if non_acgt > 0:
            m = rng.random(lens[i]) < non_acgt
            s[m] = rng.choice(np.frombuffer(b"NRYKM", dtype=np.uint8), size=int(m.sum()))
        seqs.append(s.tobytes())
    names = [b"T%06d" % i for i in range(n)]
    return names, seqs, [int(x) for x in levels]


def run_oracle(pyoracle, args, names, seqs, levels, mode, stages="all", target_hist=None):
    s = pyoracle.Sim(args.to_params(), mode)
    if args.raw_len_probs is not None:
        s.set_raw_target(args.raw_len_probs.lengths, args.raw_len_probs.probs)
    if target_hist is not None:
        s.set_target_hist(*target_hist)
    else:
        s.sample_target(args.req_frags)
    s.add_transcripts(seqs, levels)
    if stages == "target":
        return s
    s.fragment()
    if stages == "fragment":
        return s
    s.pcr()
    if stages == "pcr":
        return s
    s.sample()
    return s


def run_gpu(native, args, names, seqs, levels, stages="all", device=0, first=0, len_eff=None):
    h = native.Handle(device)
    h.set_model(args.to_params())
    if args.raw_len_probs is not None:
        h.set_raw_target(args.raw_len_probs.lengths, args.raw_len_probs.probs)
    h.sample_target(args.req_frags)
    h.set_first_index(first)
    h.add_transcripts(seqs, levels)
    if stages == "target":
        return h
    h.fragment()
    if stages == "fragment":
        return h
    if len_eff is not None:
        st = h.stats()
        h.set_len_eff(len_eff, st.low, st.high)
    h.pcr()
    if stages == "pcr":
        return h
    h.sample()
    return h
