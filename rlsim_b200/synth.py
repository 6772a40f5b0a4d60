"""Synthetic transcriptomes of the shapes BASELINE.json names.

200 000 transcripts, length = floor(lognormal(ln 1700, 0.75)) clipped to [200, 30000], per-transcript GC ~ Beta(8,8)
mapped to [0.30, 0.70], iid bases, expression from the default mixture of the reference's `sel` tool
(0.5:g:(5e3,0.5) + 0.5:g:(5e4,1000), tools/sel:137,193-201) rescaled to a requested number of molecules.
Vectorised: one pass over the concatenated bases.
"""
import numpy as np


def transcriptome(n, seed=20261019, mean_len=1700.0, sigma=0.75, lo=200, hi=30000, total_molecules=None):
    """-> (bases uint8[sum len], offsets uint64[n+1], levels uint64[n])"""
    rng = np.random.default_rng(seed)
    lens = np.clip(np.floor(rng.lognormal(np.log(mean_len), sigma, n)), lo, hi).astype(np.int64)
    gc = 0.30 + 0.40 * rng.beta(8, 8, n)
    comp = rng.random(n) < 0.5
    expr = np.where(comp, rng.gamma(0.5, 5e3 / 0.5, n), rng.gamma(1000.0, 5e4 / 1000.0, n))
    if total_molecules is not None:
        expr = expr * (total_molecules / expr.sum())
    levels = np.floor(expr).astype(np.uint64)
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    total = int(off[-1])
    thr = np.repeat(np.minimum(np.floor(gc * 256.0), 255).astype(np.uint8), lens)  # P(G or C) per base, 1/256 steps
    r = rng.integers(0, 256, size=total, dtype=np.uint8)
    bit = rng.integers(0, 2, size=total, dtype=np.uint8)
    is_gc = r < thr
    # A/T when not GC, C/G when GC
    bases = np.where(is_gc, np.where(bit == 0, ord("C"), ord("G")), np.where(bit == 0, ord("A"), ord("T"))).astype(np.uint8)
    return bases, off, levels


def write_fasta(path, bases, off, levels, first=0, count=None, width=0):
    """Expression-annotated FASTA as rlsim reads it (`>name$level`, input.go:82-101)."""
    n = len(levels) if count is None else count
    with open(path, "wb") as f:
        for i in range(first, first + n):
            f.write(b">T%07d$%d\n" % (i, int(levels[i])))
            f.write(bases[int(off[i]):int(off[i + 1])].tobytes())
            f.write(b"\n")


def fasta_text(bases, off, levels, width=70):
    """The same records as one in-memory FASTA text (uint8 array), sequences wrapped at `width` columns."""
    parts = []
    for i in range(len(levels)):
        parts.append(b">T%07d$%d\n" % (i, int(levels[i])))
        seq = bases[int(off[i]):int(off[i + 1])].tobytes()
        if width:
            parts.append(b"\n".join(seq[j:j + width] for j in range(0, len(seq), width)))
        else:
            parts.append(seq)
        parts.append(b"\n")
    return np.frombuffer(b"".join(parts), dtype=np.uint8)


def names_for(n, first=0):
    return [b"T%07d" % i for i in range(first, first + n)]
