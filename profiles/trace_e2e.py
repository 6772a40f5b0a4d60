"""One end-to-end step of the C3 workload with the library's own trace (RLSIM_TRACE=1): when each group of chunks arrived,
how long its stages took on the host's clock.  python profiles/trace_e2e.py [scale]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from rlsim_b200 import synth  # noqa: E402
from rlsim_b200.pool import GpuPool  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
n_tr = max(50, int(bench.C3_TRANSCRIPTS * scale))
b, o, l = synth.transcriptome(n_tr, seed=bench.SEED, total_molecules=bench.C3_MOLECULES * scale)
pb, po, pl = torch.from_numpy(b).pin_memory(), torch.from_numpy(o.view(np.int64)).pin_memory(), torch.from_numpy(l.view(np.int64)).pin_memory()
pool = GpuPool(0, 0, 1, None)
h = pool.handle
for step in range(4):
    if step == 3:
        os.environ["RLSIM_TRACE"] = "1"
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    pool.reset().configure(bench.make_args(int(bench.C3_FRAGS * scale), step + 1))
    h.set_first_index(0)
    t.append(time.perf_counter())
    h.add_transcripts_raw(pb, po, pl, n_tr)
    t.append(time.perf_counter())
    h.build_library()
    t.append(time.perf_counter())
    h.sample()
    t.append(time.perf_counter())
    packed, _ = h.sampled_packed()
    t.append(time.perf_counter())
    print("step %d: configure %.3f add %.3f build %.3f sample %.3f d2h %.3f ms" % ((step,) + tuple(1000 * (y - x) for x, y in zip(t[:-1], t[1:]))), file=sys.stderr)

# ---- how long do the same bytes take as bare chunked copies, and as add_transcripts without the eager groups?
os.environ.pop("RLSIM_TRACE", None)
dst = torch.empty(pb.numel(), dtype=torch.uint8, device="cuda:0")
CH = 32 << 20
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for a in range(0, pb.numel(), CH):
        dst[a:a + CH].copy_(pb[a:a + CH], non_blocking=True)
    torch.cuda.synchronize()
    print("bare chunked H2D of %d MB: %.3f ms" % (pb.numel() >> 20, 1000 * (time.perf_counter() - t0)), file=sys.stderr)
os.environ["RLSIM_NO_EAGER"] = "1"
for rep in range(3):
    torch.cuda.synchronize()
    pool.reset().configure(bench.make_args(int(bench.C3_FRAGS * scale), 50 + rep))
    h.set_first_index(0)
    t0 = time.perf_counter()
    h.add_transcripts_raw(pb, po, pl, n_tr)
    torch.cuda.synchronize()
    print("add_transcripts without eager groups (copy + ingest + prefix only): %.3f ms" % (1000 * (time.perf_counter() - t0)), file=sys.stderr)
