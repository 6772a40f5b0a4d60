"""Randomised parity: the CUDA path against the oracle's philox mode over seeded random models and transcriptomes."""
import numpy as np
import pytest

import util
from test_gpu_parity import _compare

pytestmark = pytest.mark.gpu

METHODS = ["after_prim", "after_prim_double", "after_noprim", "after_noprim_double", "pre_prim", "prim_jump"]


def _random_case(seed):
    rng = np.random.default_rng(seed)
    argv = ["-n", str(int(rng.integers(200, 20000))), "-c", str(int(rng.integers(0, 18))), "-k", str(int(rng.integers(2, 10))),
            "-p", "%.3f" % rng.uniform(1.0, 20.0), "-b", "%.2f" % rng.uniform(0, 1)]
    m = METHODS[int(rng.integers(0, 6))]
    if rng.random() < 0.5 or m == "pre_prim":
        m += ":%d" % int(rng.integers(80, 1500))
    argv += ["-f", m]
    kind = rng.integers(0, 4)
    if kind == 0:
        argv += ["-d", "1:n:(%d,%d,%d,%d)" % (rng.integers(100, 400), rng.integers(5, 60), rng.integers(30, 90), rng.integers(420, 900))]
    elif kind == 1:
        argv += ["-d", "0.7:sn:(%d,%d,%.2f,60,500) + 0.3:g:(%d,%.2f,40,700)" % (rng.integers(150, 300), rng.integers(10, 50), rng.uniform(-4, 4),
                                                                                 rng.integers(100, 300), rng.uniform(0.4, 6))]
    elif kind == 2:
        argv += ["-d", "1:g:(%d,%.2f,%d,%d)" % (rng.integers(80, 300), rng.uniform(0.3, 8), rng.integers(20, 60), rng.integers(300, 600))]
    if rng.random() < 0.4:
        argv += ["-a", "0.6:n:(%d,%d,0,%d) + 0.4:g:(%d,%.2f,0,%d)" % (rng.integers(20, 150), rng.integers(5, 40), rng.integers(150, 260),
                                                                      rng.integers(10, 120), rng.uniform(0.5, 3), rng.integers(100, 260))]
    if rng.random() < 0.3:
        argv += ["-flg", "%.2f" % rng.uniform(0, 0.8)]
    eff = rng.integers(0, 3)
    if eff == 0:
        argv += ["-eg", "(%.2f,%.2f,%.2f)" % (rng.choice([0.5, 1.0, 2.0, 3.3]), rng.uniform(0, 0.5), rng.uniform(0.5, 1.0)),
                 "-el", "(%.2f,%.2f,%.2f)" % (rng.choice([0.0, 0.1, 1.0, 2.5]), rng.uniform(0.2, 0.7), rng.uniform(0.7, 1.0))]
    elif eff == 1:
        argv += ["-e", "%.3f" % rng.uniform(0.05, 1.0)]
    if rng.random() < 0.5:
        argv += ["-sp", str(int(rng.integers(1, 1000)))]
    if rng.random() < 0.5:
        argv += ["-ss", str(int(rng.integers(1, 1000)))]
    n_tr = int(rng.integers(1, 40))
    names, seqs, levels = util.synth_transcriptome(n_tr, seed=seed + 1000, mean_len=float(rng.integers(300, 3000)), sigma=0.6, lo=1, hi=12000,
                                                   total_molecules=float(rng.integers(200, 20000)),
                                                   non_acgt=float(rng.choice([0.0, 0.0, 0.003])))
    return argv, (names, seqs, levels)


@pytest.mark.parametrize("seed", range(24))
def test_random_models(oracle, gpu, seed):
    argv, tr = _random_case(seed)
    _compare(oracle, gpu, argv, seqs_override=tr, seed=("-si", str(seed + 1)))
