"""FASTA input, unchanged format.

Host-side mirror of /root/reference/src/fasta.go:113-212 (record reader: a record starts at every '>'
once the buffer holds more than one byte; name = first line; sequence = rest with ' ', TAB, CR, LF removed,
upper-cased) and input.go:82-101,103-124 (name$level parsing, -m multiplier, per-record statistics).
Vectorised with numpy instead of the reference's byte-at-a-time reads.
"""
import re
import sys
from typing import List, Tuple

import numpy as np

_LEVEL = re.compile(rb"^[ \t\r\n]*([+-]?\d+)")
_DESPACE = bytes.maketrans(b"", b"")


def read_records(data: bytes) -> List[Tuple[bytes, bytes]]:
    """-> [(name, SEQUENCE)] in file order."""
    out = []
    if not data:
        return out
    arr = np.frombuffer(data, dtype=np.uint8)
    starts = np.flatnonzero(arr == ord(">"))
    # fasta.go:138: a '>' only opens a new record when the buffer already holds > 1 byte
    keep = [int(s) for s in starts]
    bounds = []
    prev = 0
    for s in keep:
        if s - prev > 1:
            bounds.append(s)
            prev = s
        elif s == 0:
            prev = 0
    edges = [0] + bounds + [len(data)]
    for b, e in zip(edges[:-1], edges[1:]):
        rec = data[b:e]
        if not rec:
            continue
        if rec[0:1] != b">":
            raise ValueError("Illegal fasta record:\n%s" % rec[:80].decode("latin1"))
        nl = rec.find(b"\n")
        if nl < 0:
            name, seq = rec[1:], b""
        else:
            name, seq = rec[1:nl], rec[nl + 1:]
        seq = seq.translate(_DESPACE, b" \t\r\n").upper()
        out.append((name, seq))
    return out


def parse_seq_name(s: bytes):  # input.go:82-101
    spl = s.split(b"$")
    if len(spl) != 2 or len(spl[0]) == 0:
        return None
    m = _LEVEL.match(spl[1])
    if not m:
        return None
    level = int(m.group(1))
    if level < 0:
        return None
    return spl[0], level


def load_transcripts(files: List[str], expr_mul: float = 1.0, log=None):
    """input.go:103-124 -> (names, seqs, levels); malformed names are skipped with a log line."""
    if files:
        data = b"".join(open(f, "rb").read() for f in files)  # io.MultiReader: plain concatenation
    else:
        data = sys.stdin.buffer.read()
    names, seqs, levels = [], [], []
    for name, seq in read_records(data):
        parsed = parse_seq_name(name)
        if parsed is None:
            if log:
                log("Skipping transcript with malformed name: %s" % name.decode("latin1"))
            continue
        nm, level = parsed
        level = int(float(level) * expr_mul)  # input.go:113
        names.append(nm)
        seqs.append(seq)
        levels.append(level)
    return names, seqs, levels
