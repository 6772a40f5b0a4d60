"""How does rlsim_add_transcripts (the chunk pipeline) depend on the chunk size, the form of the pipeline and the input
format?  C3 at full scale; the best of 4 timed calls per setting, host clock, milliseconds.
python profiles/trace_chunks.py"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from rlsim_b200 import _native, synth  # noqa: E402
from rlsim_b200.pool import GpuPool  # noqa: E402

n_tr = bench.C3_TRANSCRIPTS
b, o, l = synth.transcriptome(n_tr, seed=bench.SEED, total_molecules=bench.C3_MOLECULES)
pin = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).pin_memory()
pb, po, pl = pin(b), pin(o), pin(l)
codes, _ = _native.pack_bases(b)
pc = pin(codes)
pool = GpuPool(0, 0, 1, None)
h = pool.handle


def run(label, env, packed=False, reps=5):
    for k, v in env.items():
        os.environ[k] = v
    best = []
    for rep in range(reps):
        torch.cuda.synchronize()
        pool.reset().configure(bench.make_args(bench.C3_FRAGS, 100 + rep))
        h.set_first_index(0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if packed:
            h.add_transcripts_packed(pc, None, po, pl, n_tr)
        else:
            h.add_transcripts_raw(pb, po, pl, n_tr)
        t1 = time.perf_counter()
        h.build_library()
        t2 = time.perf_counter()
        best.append((1000 * (t1 - t0), 1000 * (t2 - t1)))
    for k in env:
        os.environ.pop(k, None)
    a = min(x for x, _ in best[1:])
    bl = min(y for _, y in best[1:])
    print("%-44s add %.2f build %.2f ms  eager=%d groups=%d" % (label, a, bl, bool(int(h.stats().flags) & _native.F_EAGER),
                                                                (int(h.stats().flags) >> 8) & 0xff), file=sys.stderr)


for mb in ("16", "32", "64", "128"):
    run("async ascii chunk %s MB" % mb, {"RLSIM_CHUNK_MB": mb})
for mb in ("16", "32", "64", "128"):
    run("async packed chunk %s M bases" % mb, {"RLSIM_CHUNK_MB": mb}, packed=True)
run("async ascii 32 MB, long transcripts by CTAs", {"RLSIM_PREFIX_LONG": "1"})
run("async packed 32 M, long transcripts by CTAs", {"RLSIM_PREFIX_LONG": "1"}, packed=True)
run("sync (grouped) ascii 32 MB", {"RLSIM_EAGER_SYNC": "1"})
run("sync (grouped) packed 32 M bases", {"RLSIM_EAGER_SYNC": "1"}, packed=True)
run("no eager ascii (copy+ingest+prefix)", {"RLSIM_NO_EAGER": "1"})
run("no eager packed", {"RLSIM_NO_EAGER": "1"}, packed=True)
