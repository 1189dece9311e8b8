"""Time the reference's own GPU path (oracle/_ref/ref_gpu_run, sm_100 recompile) next to this
library on the same shape: BASELINE.json configs[1] (point_mass2d, K=1e4, T=200)."""
import json
import os
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_vs_reference_gpu import run_reference_gpu  # noqa: E402
import mppi_gpu_b200 as m  # noqa: E402

K, T, A = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (10000, 200, 2)
goal, w = {1: ([1, 0], [1, 5]), 2: ([1, 0, 0, 0], [1, 1, 50, 50]),
           3: ([1, .5, .75, 0, 0, 0], [1, 1, 1, 5, 5, 5])}[A]
x0, U0 = np.zeros(2 * A, np.float32), np.zeros((T, A), np.float32)
with tempfile.TemporaryDirectory() as d:
    ref = run_reference_gpu(Path(d), K, T, A, 0.1, x0, U0, goal, w, 6)
ref_ms = float(np.median(ref["ms"][1:]))
ctl = m.PointMassModel(K, T, 0.1, 2 * A, A)
ctl.memcpy_set_data(x0, U0, goal, w)
import time
for _ in range(20):
    ctl.get_act()
lat = []
for _ in range(200):
    t0 = time.perf_counter()
    ctl.get_act()
    lat.append((time.perf_counter() - t0) * 1e3)
ours_ms = float(np.median(lat))
print(json.dumps({"shape": {"K": K, "T": T, "A": A},
                  "reference_gpu_path_sm100_recompile_ms_per_get_act": ref_ms,
                  "reference_gpu_rollout_steps_per_s": K * T / (ref_ms * 1e-3),
                  "this_library_ms_per_get_act_p50": ours_ms,
                  "this_library_rollout_steps_per_s": K * T / (ours_ms * 1e-3),
                  "speedup": ref_ms / ours_ms}))
