"""Extract the method sets of the reference's Go interfaces the GPU shim implements (Pooler: src/pool.go:36-48, Sampler:
src/sampler.go:39-42) into tests/golden/go_interfaces.json.  Run in the build container, where /root/reference exists:

    python tests/golden/make_go_interfaces.py

tests/test_host.py checks go/gpu_pool.go against the committed file (and, where the reference is mounted, that the
file is still what the reference declares)."""
import json
import os
import re

REF = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))


def interface_methods(path, name):
    src = open(path).read()
    m = re.search(r"type\s+%s\s+interface\s*\{(.*?)\n\}" % name, src, re.S)
    out = {}
    for line in m.group(1).strip().splitlines():
        line = line.strip()
        mm = re.match(r"(\w+)\((.*?)\)\s*(.*)$", line)
        if not mm:
            continue
        params = [p.strip() for p in mm.group(2).split(",") if p.strip()]
        out[mm.group(1)] = {"n_params": len(params), "params": mm.group(2).strip(), "returns": mm.group(3).strip()}
    return out


def extract():
    return {"Pooler": interface_methods(os.path.join(REF, "pool.go"), "Pooler"),
            "Sampler": interface_methods(os.path.join(REF, "sampler.go"), "Sampler")}


if __name__ == "__main__":
    json.dump(extract(), open(os.path.join(HERE, "go_interfaces.json"), "w"), indent=1, sort_keys=True)
