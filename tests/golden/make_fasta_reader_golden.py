#!/usr/bin/env python3
"""Generate tests/golden/fasta_reader.json: what the reference's release binary makes of unusual FASTA texts.

Run in the build container (oracle/_ref holds the unpacked release):  python tests/golden/make_fasta_reader_golden.py
For every text the binary runs `rlsim -g -t 1 -n 200 -si 3 [-m x]`; the fixture keeps the two report panels the reader
feeds (Transcript lengths: length -> sum of levels; Expression levels: level -> count), the "Skipping transcript with
malformed name" log lines and the transcript names seen in the FASTA output.  These pin oracle/fasta_reader.py
(fasta.go:113-212, input.go:82-124).  Not pinned: a record without any newline (the binary panics on Record[1:-1]).
"""
import json
import os
import re
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import refbin  # noqa: E402
import util  # noqa: E402

base = open(util.BASIC_FASTA, "rb").read()
recs = [r for r in base.split(b">") if r]
seqs = [b"".join(r.split(b"\n")[1:]) for r in recs]   # the five fixture sequences, one line each


def wrap(s, w):
    return b"\n".join(s[i:i + w] for i in range(0, len(s), w))


TEXTS = {
    "crlf_lowercase": (b">a$900\r\n" + wrap(seqs[0], 60).replace(b"\n", b"\r\n") + b"\r\n>b$700\r\n" + wrap(seqs[1].lower(), 70).replace(b"\n", b"\r\n") + b"\r\n", []),
    "spaces_blank_lines": (b">a$500 \t\n" + wrap(seqs[0], 50).replace(b"\n", b" \t\n\n") + b"\n\n\n>b two words$ 300\n" + wrap(seqs[2], 80) + b"\n", []),
    "gt_inside_line": (b">a$400\n" + seqs[0][:600] + b">b$350\n" + wrap(seqs[1], 61) + b"\n", []),
    "double_gt": (b">>a$450\n" + wrap(seqs[0], 60) + b"\n>>b$250\n" + wrap(seqs[3], 60) + b"\n", []),   # (">>>" makes a record ">>" without newline: the binary panics)
    "malformed_names": (b">ok$600\n" + wrap(seqs[0], 60) + b"\n>no_dollar\n" + seqs[1][:300] + b"\n>x$1$2\n" + seqs[1][:300] + b"\n>$7\n" + seqs[1][:300] +
                        b"\n>neg$-5\n" + seqs[1][:300] + b"\n>plus$+5\n" + seqs[1][:300] + b"\n>word$abc\n" + seqs[1][:300] +
                        b"\n>huge$18446744073709551616\n" + seqs[1][:300] + b"\n>junk$12abc\n" + wrap(seqs[2], 60) + b"\n>lead$   34\n" + wrap(seqs[3], 60) + b"\n", []),
    "levels_and_multiplier": (b">a$0007\n" + wrap(seqs[0], 60) + b"\n>b$1001\n" + wrap(seqs[1], 60) + b"\n>zero$0\n" + wrap(seqs[2], 60) + b"\n", ["-m", "0.5"]),
    "empty_sequence_record": (b">a$800\n" + wrap(seqs[0], 60) + b"\n>empty$9\n>b$100\n" + wrap(seqs[4], 60) + b"\n", []),
    "no_final_newline": (b">a$800\n" + wrap(seqs[0], 60) + b"\n>b$120\n" + seqs[4], []),
}

out = {}
for name, (text, extra) in TEXTS.items():
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "in.fas")
        open(fa, "wb").write(text)
        r = refbin.run(["-n", "200", "-si", "3"] + extra, fa)
    skipped = re.findall(r"Skipping transcript with malformed name: (.*)", r["log"])
    names = sorted({m.group(1).decode("latin1") for m in re.finditer(rb"^>Fg_\d+_(.*) \(Strand [+-] Offset \d+ -- \d+\)$", r["fasta"], re.M)})
    out[name] = {"text": text.decode("latin1"), "argv": extra, "skipped": skipped, "out_names": names,
                 "panels": {p: r["report"][p]["Data"] for p in ("Transcript lengths", "Expression levels")}}
    print(name, out[name]["panels"], skipped, names)
with open(os.path.join(util.GOLDEN, "fasta_reader.json"), "w") as f:
    json.dump(out, f, sort_keys=True)
