"""BASELINE.json configs[3]: K sweep 1e3..1e7 at T=200 for point_mass1d/2d/3d, per-kernel HBM
roofline fraction.  Writes one JSON object per (A, K) to stdout."""
import json
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import mppi_gpu_b200 as m  # noqa: E402
from mppi_gpu_b200 import capi  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
CFG = {1: ([1, 0], [1, 5]), 2: ([1, 0, 0, 0], [1, 1, 50, 50]),
       3: ([1, .5, .75, 0, 0, 0], [1, 1, 1, 5, 5, 5])}
T = 200
for A in (1, 2, 3):
    for K in (1000, 10000, 100000, 1000000, 10000000):
        steps = 200 if K <= 100000 else (30 if K <= 1000000 else 8)
        best = None
        for flags in (0, capi.FLAG_FUSED_SAMPLING, capi.FLAG_STEP_KERNEL, capi.FLAG_AUTO_CHAIN):
            if flags == capi.FLAG_STEP_KERNEL and K < 100000:
                continue
            ctl = m.PointMassModel(K, T, 0.1, 2 * A, A, flags=flags)
            resolved = ctl.flags()
            ctl.memcpy_set_data(np.zeros(2 * A), np.zeros(T * A), *CFG[A])
            for _ in range(5):
                ctl.get_act()
            ctl.timer_start()
            for _ in range(steps):
                ctl.step_enqueue()
            ms = ctl.timer_stop() / steps
            ctl.step_wait()
            ctl.set_profiling(True)
            for _ in range(steps):
                ctl.get_act()
            kt = {k: t / n for k, (t, n) in ctl.kernel_times().items() if n}
            ctl.close()
            eps_bytes = 4.0 * K * T * A
            row = {"A": A, "K": K, "T": T, "flags": flags, "resolved_flags": resolved, "ms_per_step": ms,
                   "rollout_steps_per_s": K * T / (ms * 1e-3),
                   "step_hbm_frac_of_3_pass_roofline": (3 * eps_bytes + 16.0 * K) / (ms * 1e-3) / 1e9 / PEAK,
                   "step_hbm_frac_of_2_pass_roofline": (2 * eps_bytes + 8.0 * K) / (ms * 1e-3) / 1e9 / PEAK,
                   "kernels": {k: {"ms": v, "hbm_frac": (eps_bytes / (v * 1e-3) / 1e9 / PEAK)
                                   if k in ("sample", "rollout", "average") else None}
                               for k, v in kt.items()}}
            if best is None or ms < best["ms_per_step"]:
                best = row
            print(json.dumps(row), flush=True)
