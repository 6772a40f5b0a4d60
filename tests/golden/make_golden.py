#!/usr/bin/env python3
"""Generate tests/golden/*.json from the reference's own release binary.

Run in the build container (where oracle/_ref holds releases/rlsim-1.4_amd64.tar.gz unpacked):
    python tests/golden/make_golden.py
For every configuration of tests/util.py:PIN_CONFIGS it runs
    rlsim -g -t 1 -si <seed> <flags> tests/golden/test_transcripts.fas
and stores the report panels that the binary reproduces run-to-run for a fixed -si
(SURVEY.md section 8c): target lengths, poly-A tail lengths, fragment lengths after fragmentation, the two
efficiency functions and the input statistics.  These are the oracle's pins.
test_transcripts.fas and raw_params.json are the reference's own fixtures (src/test/basic/).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import refbin  # noqa: E402
import util  # noqa: E402

for name, argv in util.PIN_CONFIGS.items():
    seed = util.pin_seed(name)
    r = refbin.run(list(argv) + ["-si", str(seed)], util.BASIC_FASTA, keep_fasta=False)
    out = {"argv": argv, "seed": seed, "panels": {}}
    for p in util.PIN_PANELS:
        if p in r["report"]:
            out["panels"][p] = r["report"][p]["Data"]
    with open(os.path.join(util.GOLDEN, name + ".json"), "w") as f:
        json.dump(out, f, sort_keys=True)
    print(name, {k: len(v) for k, v in out["panels"].items()})
