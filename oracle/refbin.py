"""Run the reference's own release binary (oracle/_ref/rlsim-1.4_amd64/bin/rlsim) and parse what it emits.

TEST INFRASTRUCTURE ONLY.  The binary is a black box: CLI in -> FASTA on stdout + JSON report.
"""
import json
import os
import re
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(HERE, "_ref", "rlsim-1.4_amd64", "bin", "rlsim")
_HDR = re.compile(rb"^>Fg_(\d+)_(\S+) \(Strand ([+-]) Offset (\d+) -- (\d+)\)$")


def available():
    return os.access(BIN, os.X_OK)


def run(args, fasta_path, threads=1, keep_fasta=True, timeout=3600):
    """-> dict(report=..., fasta=bytes|None, log=str, wall=seconds)."""
    import time
    with tempfile.TemporaryDirectory() as td:
        rep = os.path.join(td, "rep.json")
        cmd = [BIN, "-g", "-v", "-t", str(threads), "-r", rep] + list(args) + [fasta_path]
        t0 = time.time()
        p = subprocess.run(cmd, stdout=subprocess.PIPE if keep_fasta else subprocess.DEVNULL,
                           stderr=subprocess.PIPE, cwd=td, timeout=timeout)
        wall = time.time() - t0
        if p.returncode != 0:
            raise RuntimeError("reference binary failed: %s" % p.stderr.decode()[-2000:])
        report = json.load(open(rep))
    return dict(report=report, fasta=p.stdout if keep_fasta else None, log=p.stderr.decode(), wall=wall)


def panel(report, title):
    """JSON panel -> {int key: value} (float keys for the GC panel)."""
    d = report[title]["Data"]
    out = {}
    for k, v in d.items():
        try:
            out[int(k)] = v
        except ValueError:
            out[float(k)] = v
    return out


def parse_fasta(data):
    """-> list of (id, name, strand, start, end, seq)."""
    out = []
    lines = data.split(b"\n")
    for i in range(0, len(lines) - 1, 2):
        m = _HDR.match(lines[i])
        if not m:
            raise ValueError("bad header %r" % lines[i])
        out.append((int(m.group(1)), m.group(2), m.group(3).decode(), int(m.group(4)), int(m.group(5)), lines[i + 1]))
    return out


def stage_times(log):
    """Per-stage wall times from the -v log timestamps."""
    ts = {}
    for line in log.splitlines():
        m = re.match(r"^(\d+):(\d+):(\d+\.\d+) (.*)$", line)
        if not m:
            continue
        t = int(m.group(1)) * 3600 + int(m.group(2)) * 60 + float(m.group(3))
        msg = m.group(4)
        for key, pat in (("start", "Starting up"), ("target_done", "Finished sampling target lengths"),
                         ("frag_start", "Fragmenting transcripts"), ("frag_done", "Initialized"),
                         ("sample_done", "Total number of fragments in the pool")):
            if msg.startswith(pat) and key not in ts:
                ts[key] = t
    return ts
