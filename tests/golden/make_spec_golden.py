#!/usr/bin/env python3
"""Freeze the `--rng=philox` specification: tests/golden/philox_spec.json holds, for every pinned configuration, digests of
what the normative restatement (oracle/oracle_px.c) produces - species table, amplified counts, sampled records, FASTA
text.  The specification is this repository's own, so the fixture is not a pin against the
reference; it is a tripwire: a change that moves the oracle and the CUDA path together (they are compared with each other
everywhere else) cannot go unnoticed.  Regenerate ONLY when the specification is changed on purpose.
    python tests/golden/make_spec_golden.py
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from oracle import pyoracle  # noqa: E402
import util  # noqa: E402


def digest(sp, o, names):
    """sp = o.species() taken after PCR and BEFORE the draw (the oracle decrements the counts while it samples)."""
    sm = o.sampled()
    h = lambda *arrs: hashlib.sha256(b"".join(np.ascontiguousarray(a).astype("<u8").tobytes() for a in arrs)).hexdigest()
    return {"n_species": int(len(sp["tr"])), "n_sampled": int(len(sm["tr"])), "pool_total": int(o.pool_total),
            "species": h(sp["tr"], sp["start"], sp["end"], sp["count0"]), "counts": h(sp["count"]),
            "sampled": h(sm["tr"], sm["start"], sm["end"], sm["minus"], sm["id"]),
            "fasta": hashlib.sha256(o.fasta(names)).hexdigest()}


if __name__ == "__main__":
    pyoracle.build()
    out = {}
    for name, argv in util.PIN_CONFIGS.items():
        if name == "c2":
            argv = [a if a != "500000" else "20000" for a in argv]  # keep the CPU suite short
        args = util.parse(list(argv) + ["-si", str(util.pin_seed(name))])
        names, seqs, levels = util.transcripts(args)
        o = util.run_oracle(pyoracle, args, names, seqs, levels, pyoracle.MODE_PHILOX, stages="pcr")
        sp = o.species()
        o.sample()
        out[name] = {"argv": list(argv), "seed": util.pin_seed(name), **digest(sp, o, names)}
        print(name, out[name]["n_species"], out[name]["n_sampled"], out[name]["fasta"][:12])
    json.dump(out, open(os.path.join(util.GOLDEN, "philox_spec.json"), "w"), indent=1, sort_keys=True)
