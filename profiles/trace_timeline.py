"""Device-side timeline of the chunk pipeline (RLSIM_TRACE=2: timing events at the stage boundaries of every chunk), C3 at
full scale, packed and ASCII input.  python profiles/trace_timeline.py"""
import os
import sys

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
pc = pin(_native.pack_bases(b)[0])
pool = GpuPool(0, 0, 1, None)
h = pool.handle
for packed in (True, False):
    for rep in range(4):
        if rep == 3:
            os.environ["RLSIM_TRACE"] = "2"
            print("==== %s input" % ("packed" if packed else "ascii"), file=sys.stderr)
        torch.cuda.synchronize()
        pool.reset().configure(bench.make_args(bench.C3_FRAGS, 200 + rep))
        h.set_first_index(0)
        if packed:
            h.add_transcripts_packed(pc, None, po, pl, n_tr)
        else:
            h.add_transcripts_raw(pb, po, pl, n_tr)
        h.build_library()
        os.environ.pop("RLSIM_TRACE", None)
