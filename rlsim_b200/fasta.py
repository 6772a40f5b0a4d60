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

_LEVEL = re.compile(rb"^[ \t\r\n\v\f]*(\d+)")  # TrimSpace + Sscanf("%d") into a uint64: no sign
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
    if level >= 1 << 64:  # strconv.ParseUint: value out of range
        return None
    return spl[0], level


def load_transcripts(files: List[str], expr_mul: float = 1.0, log=None, data: bytes = None):
    """input.go:103-124 -> (names, seqs, levels); malformed names are skipped with a log line."""
    if data is None:
        data = read_input(files)
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


def read_input(files: List[str]) -> bytes:
    """fasta.go:98-104 / input.go:45-60: the files concatenated (io.MultiReader), or stdin."""
    if files:
        return b"".join(open(f, "rb").read() for f in files)
    return sys.stdin.buffer.read()


def load_transcripts_device(handle, data: bytes, expr_mul: float = 1.0, log=None):
    """Device-side reader (rlsim_add_fasta): the text is split, despaced, upper-cased and its `name$level` lines
    parsed on the GPU; the records are added to `handle`.  -> (names, seq_lens, levels) of the added transcripts.
    Same records, same skips and same log lines as load_transcripts + add_transcripts."""
    idx = handle.add_fasta(data, expr_mul)
    names, lens, levels = [], [], []
    start, line_len, name_len = idx["name_start"], idx["line_len"], idx["name_len"]
    for r in range(len(start)):
        s0 = int(start[r])
        if not idx["valid"][r]:
            if log:
                log("Skipping transcript with malformed name: %s" % data[s0:s0 + int(line_len[r])].decode("latin1"))
            continue
        names.append(bytes(data[s0:s0 + int(name_len[r])]))
        lens.append(int(idx["seq_len"][r]))
        levels.append(int(idx["level"][r]))
    return names, lens, levels
