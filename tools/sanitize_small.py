"""Small end-to-end exercise of every kernel variant (for compute-sanitizer)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import mppi_gpu_b200 as m  # noqa: E402
from mppi_gpu_b200 import capi  # noqa: E402

CFG = {1: ([1, 0], [1, 5]), 2: ([1, 0, 0, 0], [1, 1, 50, 50]),
       3: ([1, .5, .75, 0, 0, 0], [1, 1, 1, 5, 5, 5]), 4: ([1, .5, .75, -.5, 0, 0, 0, 0], [1] * 8)}
rs = np.random.RandomState(0)
n = 0
for A in (1, 2, 3, 4):
    for K, T in ((1027, 37), (5, 3), (3000, 50)):
        for flags, env in ((0, {}), (capi.FLAG_SPLIT_KERNELS | capi.FLAG_STRICT_ARITH, {}),
                           (capi.FLAG_FUSED_SAMPLING, {}), (0, {"MPPI_ROLLOUT_TMA": "0", "MPPI_ROLLOUT_SPT": "2"}),
                           (capi.FLAG_NO_GRAPH, {"MPPI_ROLLOUT_TMA": "0", "MPPI_ROLLOUT_SPT": "1"}),
                           (0, {"MPPI_ROLLOUT_TMA": "1", "MPPI_ROLLOUT_TMA_W": "256"})):
            for k, v in env.items():
                os.environ[k] = v
            ctl = m.PointMassModel(K, T, 0.1, 2 * A, A, flags=flags, seed=7)
            for k in env:
                del os.environ[k]
            ctl.memcpy_set_data(np.zeros(2 * A), 0.1 * rs.standard_normal(T * A), *CFG[A])
            ctl.get_act()
            ctl.set_x(0.01 * rs.standard_normal(2 * A))
            ctl.get_act()
            ctl.set_noise((0.1 * rs.standard_normal((K, T, A))).astype(np.float32))
            ctl.get_act()
            inf = ctl.get_inf(want_x=True)
            assert np.isfinite(inf["cost"]).all() and np.isfinite(inf["u"]).all()
            ctl.set_profiling(True)
            ctl.get_act()
            ctl.close()
            n += 1
print("SANITIZE_SCRIPT_OK", n, "controllers")
