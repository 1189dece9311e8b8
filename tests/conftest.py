import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["cfgtest", "testcu", "pm1d", "pm2d", "pm3d", "pm3d_ragged", "pm2d_damped",
                "pm3d_gains"]


def golden_gains(g):
    """(state_gain, act_gain) of a golden case made with caller-given gains, else None."""
    return (g["state_gain"], g["act_gain"]) if "state_gain" in g else None

# reference configs (config/point_mass{1,2,3}d.yaml): goal and cost.w per action dim
REF_CFG = {
    1: dict(goal=[1, 0], w=[1, 5]),
    2: dict(goal=[1, 0, 0, 0], w=[1, 1, 50, 50]),
    3: dict(goal=[1, 0.5, 0.75, 0, 0, 0], w=[1, 1, 1, 5, 5, 5]),
    4: dict(goal=[1, 0.5, 0.75, -0.5, 0, 0, 0, 0], w=[1, 1, 1, 2, 5, 5, 5, 3]),
}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: d[k] for k in d.files}


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle as po
    po.lib()
    return po


def make_inputs(K, T, A, seed, sigma=0.025, u_scale=0.1, x_scale=0.05):
    rs = np.random.RandomState(seed)
    eps = (sigma * rs.standard_normal((K, T, A))).astype(np.float32)
    U = (u_scale * rs.standard_normal((T, A))).astype(np.float32)
    x0 = (x_scale * rs.standard_normal(2 * A)).astype(np.float32)
    return x0, U, eps


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
