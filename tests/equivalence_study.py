#!/usr/bin/env python3
"""Config C4 of BASELINE.json in miniature: PCR + sampling pseudo-replicates (20 seeds) of the reference's release
binary against 20 seeds of the CUDA path on a C3-shaped synthetic transcriptome, KS / chi-square per statistic.

    python tests/equivalence_study.py [n_transcripts] > profiles/r1_equivalence_study.md      (on the GPU box)

Not a pytest (runs ~2 minutes); the asserting version on the reference's 5-transcript fixture is
tests/test_stat_equivalence.py.  The binary and the oracle are used only as the checker.
"""
import os
import sys
import tempfile

import numpy as np
from scipy import stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from oracle import refbin  # noqa: E402
from rlsim_b200 import _native, synth  # noqa: E402
from test_stat_equivalence import _chi2, _expand  # noqa: E402

N_TR = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
N_SEEDS = 20
FLAGS = ["-d", "0.9:sn:(600,50,4,300,2000) + 0.1:n:(700,1,600,2000)", "-f", "after_prim_double", "-c", "15", "-eg", "(1.0,0.5,0.8)",
         "-el", "(0.1,0.7,1.0)"]
N_FRAGS = 100 * N_TR


def gc_percent(seq):
    a = np.frombuffer(seq, dtype=np.uint8)
    return 100.0 * np.count_nonzero((a == ord("G")) | (a == ord("C"))) / max(len(a), 1)


def main():
    bases, off, levels = synth.transcriptome(N_TR, seed=77, total_molecules=50 * N_TR)
    names = synth.names_for(N_TR)
    seqs = [bases[int(off[i]):int(off[i + 1])].tobytes() for i in range(N_TR)]
    top = np.argsort(-levels.astype(np.int64))[:8]
    runs = {"binary": [], "cuda": []}
    pooled = {"binary": {}, "cuda": {}}
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "t.fas")
        synth.write_fasta(fa, bases, off, levels)
        for s in range(1, N_SEEDS + 1):
            r = refbin.run(["-n", str(N_FRAGS), "-si", str(s)] + FLAGS, fa, keep_fasta=True)
            rep = r["report"]
            recs = refbin.parse_fasta(r["fasta"])
            nf = sum(rep["Fragdist after fragmentation"]["Data"].values())
            per_tr = np.zeros(N_TR)
            for _, name, _, _, _, _ in recs:
                per_tr[int(name[1:])] += 1
            runs["binary"].append(dict(amp=sum(rep["Fragdist after PCR"]["Data"].values()) / nf, nfrag=nf, sampled=len(recs),
                                       gc=float(np.mean([gc_percent(x[5]) for x in recs[::7]])), per_tr=per_tr / max(len(recs), 1),
                                       minus=np.mean([x[2] == "-" for x in recs]), mean_len=np.mean([x[4] - x[3] for x in recs])))
            for k, v in rep["Fragdist after fragmentation"]["Data"].items():
                pooled["binary"][int(k)] = pooled["binary"].get(int(k), 0) + v
    for s in range(1001, 1001 + N_SEEDS):
        args = util.parse(["-n", str(N_FRAGS), "-si", str(s)] + FLAGS)
        h = util.run_gpu(_native, args, names, seqs, [int(x) for x in levels])
        st = h.stats()
        out = h.sampled()
        fasta = h.fasta(names).split(b"\n")
        per_tr = np.bincount(out["tr"], minlength=N_TR).astype(np.float64)
        frag = h.hist_dict(_native.H_AFTER_FRAG, 2100)
        runs["cuda"].append(dict(amp=st.pool_total / st.n_fragments, nfrag=st.n_fragments, sampled=st.n_sampled,
                                 gc=float(np.mean([gc_percent(fasta[2 * i + 1]) for i in range(0, st.n_sampled, 7)])),
                                 per_tr=per_tr / max(st.n_sampled, 1), minus=float(out["minus"].mean()),
                                 mean_len=float((out["end"] - out["start"]).mean())))
        for k, v in frag.items():
            pooled["cuda"][k] = pooled["cuda"].get(k, 0) + v
    rows = []
    rows.append(("fragment length after fragmentation, pooled (chi2)", _chi2(pooled["binary"], pooled["cuda"])))
    rows.append(("fragment length after fragmentation, pooled (KS)", stats.ks_2samp(_expand(pooled["binary"]), _expand(pooled["cuda"]))[1]))
    for key, label in (("nfrag", "fragments kept per run"), ("amp", "PCR copy number: pool / fragments per run"), ("sampled", "sampled fragments per run"),
                       ("gc", "mean GC% of sampled fragments per run"), ("mean_len", "mean sampled fragment length per run"), ("minus", "minus-strand fraction per run")):
        a, b = [r[key] for r in runs["binary"]], [r[key] for r in runs["cuda"]]
        rows.append((label + " (KS)", stats.ks_2samp(a, b)[1]))
        pt = stats.ttest_ind(a, b, equal_var=False)[1] if (np.std(a) > 0 or np.std(b) > 0) else (1.0 if np.mean(a) == np.mean(b) else 0.0)
        rows.append((label + " (Welch t)", pt))
    a, b = [r["amp"] for r in runs["binary"]], [r["amp"] for r in runs["cuda"]]
    rows.append(("PCR copy number variance (Levene)", stats.levene(a, b)[1]))
    for t in top:
        a, b = [r["per_tr"][t] for r in runs["binary"]], [r["per_tr"][t] for r in runs["cuda"]]
        rows.append(("sampled fraction of transcript T%07d (level %d) per run (KS)" % (t, levels[t]), stats.ks_2samp(a, b)[1]))
    print("# Equivalence study: release binary vs CUDA path, %d seeds each\n" % N_SEEDS)
    print("Workload: %d synthetic transcripts (C3 shape), %d molecules, `-n %d %s`." % (N_TR, int(levels.sum()), N_FRAGS, " ".join(FLAGS)))
    print("Binary seeds 1..%d, CUDA seeds 1001..%d (independent streams).  Post-PCR quantities are compared as per-run statistics" % (N_SEEDS, 1000 + N_SEEDS))
    print("(20 runs against 20) because pooled counts are over-dispersed (DESIGN.md section 2).\n")
    print("| statistic | p-value |\n|---|---:|")
    for label, p in rows:
        print("| %s | %.4f |" % (label, p))
    low = [l for l, p in rows if not p > 0.01]
    print("\n%d of %d statistics at p > 0.01%s" % (len(rows) - len(low), len(rows), "" if not low else " (below: %s - expected by chance at this many tests: %.1f)" % ("; ".join(low), 0.01 * len(rows))))
    ab = np.mean([r["amp"] for r in runs["binary"]]), np.mean([r["amp"] for r in runs["cuda"]])
    print("\nMean amplification factor: binary %.2f, CUDA %.2f; mean sampled: binary %.0f, CUDA %.0f." % (ab[0], ab[1], np.mean([r["sampled"] for r in runs["binary"]]), np.mean([r["sampled"] for r in runs["cuda"]])))


if __name__ == "__main__":
    main()
