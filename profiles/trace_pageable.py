"""rlsim_add_transcripts from pageable (plain numpy) arrays: how many staging threads does the host link want?
C3 at full scale; best of 4 calls per setting.  python profiles/trace_pageable.py"""
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
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), file=sys.stderr)
for threads, nt in (("4", 1), ("6", 1), ("8", 1), ("12", 1), ("4", 0), ("6", 0), ("8", 0), ("12", 0)):
    os.environ["RLSIM_STAGE_THREADS"] = threads
    if nt:
        os.environ.pop("RLSIM_NO_STREAM_COPY", None)
    else:
        os.environ["RLSIM_NO_STREAM_COPY"] = "1"
    pool = GpuPool(0, 0, 1, None)
    h = pool.handle
    ts = []
    for rep in range(5):
        torch.cuda.synchronize()
        pool.reset().configure(bench.make_args(bench.C3_FRAGS, 300 + rep))
        h.set_first_index(0)
        t0 = time.perf_counter()
        h.add_transcripts_raw(b, o, l, n_tr)
        ts.append(1000 * (time.perf_counter() - t0))
        h.build_library()
    print("stage threads %2s, %s: add %.2f ms (staged=%d)" % (threads, "non-temporal stores" if nt else "memcpy", min(ts[1:]), bool(int(h.stats().flags) & _native.F_PAGEABLE_STAGED)), file=sys.stderr)
    h.close()
# plain memcpy rate of one thread, for scale
src = b
dst = np.empty_like(b)
t0 = time.perf_counter(); np.copyto(dst, src); t1 = time.perf_counter()
t0 = time.perf_counter(); np.copyto(dst, src); t1 = time.perf_counter()
print("one-thread numpy copy of %d MB: %.1f ms (%.1f GB/s)" % (b.size >> 20, 1000 * (t1 - t0), b.size / (t1 - t0) / 1e9), file=sys.stderr)
