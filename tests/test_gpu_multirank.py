"""The multi-rank selection on real ranks: two processes, two GPUs, NCCL (tests/multirank_check.py).

Needs at least two visible GPUs; on a single-GPU box the test is skipped (the emulated two- and three-rank tests in
test_gpu_parity.py still run there).  profiles/r2_multirank_n2.json / _n4.json hold the outcome of the same script run with
`gpurun --gpus 2` / `--gpus 4` this round."""
import json
import os
import subprocess
import sys
import tempfile

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_reproduce_one_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "out.json")
        subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                               "--master-port", "29741", os.path.join(ROOT, "tests", "multirank_check.py"), out],
                              env=dict(os.environ, MASTER_ADDR="127.0.0.1"), timeout=900)
        res = json.load(open(out))
    assert res["ok"] and res["world"] == 2
    for name, c in res["cases"].items():
        assert c["library"] == c["one_gpu"] == c["torch"], name
    assert res["cases"]["exhaustion"]["missing"] > 0
