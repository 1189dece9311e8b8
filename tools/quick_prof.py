"""Quick per-kernel timing of one configuration (development aid; bench.py is the contract)."""
import argparse
import json
import sys
import os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import mppi_gpu_b200 as m
from mppi_gpu_b200 import capi

ap = argparse.ArgumentParser()
ap.add_argument("-K", type=int, default=1000000)
ap.add_argument("-T", type=int, default=200)
ap.add_argument("-A", type=int, default=3)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--flags", type=int, default=0)
a = ap.parse_args()
cfg = {1: ([1, 0], [1, 5]), 2: ([1, 0, 0, 0], [1, 1, 50, 50]),
       3: ([1, .5, .75, 0, 0, 0], [1, 1, 1, 5, 5, 5]), 4: ([1, .5, .75, -.5, 0, 0, 0, 0], [1] * 8)}[a.A]
ctl = m.PointMassModel(a.K, a.T, 0.1, 2 * a.A, a.A, flags=a.flags, verbose=True)
ctl.memcpy_set_data(np.zeros(2 * a.A), np.zeros(a.T * a.A), cfg[0], cfg[1])
for _ in range(3):
    ctl.get_act()
ctl.timer_start()
for _ in range(a.steps):
    ctl.step_enqueue()
ms = ctl.timer_stop() / a.steps
ctl.step_wait()
ctl.set_profiling(True)
for _ in range(a.steps):
    ctl.get_act()
kt = ctl.kernel_times()
ctl.set_profiling(False)
eps_bytes = 4.0 * a.K * a.T * a.A
out = {"K": a.K, "T": a.T, "A": a.A, "flags": a.flags, "graph_ms_per_step": ms,
       "rollout_steps_per_s": a.K * a.T / (ms * 1e-3)}
for k, (t, n) in kt.items():
    if n:
        out[k + "_ms"] = t / n
for k in ("sample", "rollout", "average"):
    if k + "_ms" in out:
        out[k + "_GBs"] = eps_bytes / (out[k + "_ms"] * 1e-3) / 1e9
print(json.dumps(out))
