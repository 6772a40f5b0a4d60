"""One monolithic pass of the C3 workload (every kernel launches once at full size) for `ncu --set full`.

    python profiles/profile_run.py [scale]

The eager chunk pipeline is switched off and the transcripts arrive as one chunk, so each stage is one launch.
"""
import os
import sys

os.environ["RLSIM_NO_EAGER"] = "1"
os.environ["RLSIM_CHUNK_MB"] = "4096"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from rlsim_b200 import synth  # noqa: E402
from rlsim_b200.pool import GpuPool  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
n_tr = max(50, int(bench.C3_TRANSCRIPTS * scale))
bases_np, off_np, levels_np = synth.transcriptome(n_tr, seed=bench.SEED, total_molecules=bench.C3_MOLECULES * scale)
bases = torch.from_numpy(bases_np).pin_memory()
off = torch.from_numpy(off_np.view(np.int64)).pin_memory()
levels = torch.from_numpy(levels_np.view(np.int64)).pin_memory()
pool = GpuPool(0, 0, 1, None)
h = pool.handle
for step in range(2):  # the second pass runs with every buffer allocated
    pool.reset().configure(bench.make_args(int(bench.C3_FRAGS * scale), step + 1))
    h.set_first_index(0)
    h.add_transcripts_raw(bases, off, levels, n_tr)
    h.build_library()
    h.sample()
    torch.cuda.synchronize()
st = h.stats()
expressed = levels_np > 0
nt = int((off_np[1:] - off_np[:-1]).astype(np.int64)[expressed].sum()) + int(expressed.sum()) * int(st.polya_max)
print("nt %d molecules %d fragments %d species %d sampled %d" % (nt, st.n_molecules, st.n_fragments, st.n_species, st.n_sampled))
