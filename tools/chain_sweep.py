"""Every kernel chain on a grid of shard shapes: device time per control step (graph replays back
to back, CUDA events).  Basis of the MPPI_FLAG_AUTO_CHAIN policy (controller.cu: auto_chain) and
of BASELINE.json configs[3] (K sweep across point_mass1d/2d/3d).  One JSON object per (A, T, K)
on stdout: {"A","T","K","ms": {chain: ms}, "best", "auto", "auto_ms", "auto_vs_best"}."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import mppi_gpu_b200 as m  # noqa: E402
from mppi_gpu_b200 import capi  # noqa: E402

CFG = {1: ([1, 0], [1, 5]), 2: ([1, 0, 0, 0], [1, 1, 50, 50]),
       3: ([1, .5, .75, 0, 0, 0], [1, 1, 1, 5, 5, 5]),
       4: ([1, .5, .75, -.5, 0, 0, 0, 0], [1, 1, 1, 2, 5, 5, 5, 3])}
CHAINS = {"unfused": 0, "fused": capi.FLAG_FUSED_SAMPLING, "step": capi.FLAG_STEP_KERNEL,
          "tile": capi.FLAG_TILE_KERNEL}


def time_chain(K, T, A, flags, steps):
    ctl = m.PointMassModel(K, T, 0.1, 2 * A, A, flags=flags)
    ctl.memcpy_set_data(np.zeros(2 * A), np.zeros(T * A), *CFG[A])
    for _ in range(5):
        ctl.get_act()
    ctl.timer_start()
    for _ in range(steps):
        ctl.step_enqueue()
    ms = ctl.timer_stop() / steps
    ctl.step_wait()
    n0 = ctl.launch_count()
    ctl.get_act()
    per_step = ctl.launch_count() - n0
    got = ctl.flags()
    ctl.close()
    return ms, got, per_step


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="3:200,2:200,1:200,2:50,4:100")
    ap.add_argument("--K", default="1000,10000,30000,100000,125000,250000,500000,1000000")
    ap.add_argument("--chains", default="unfused,fused,step,tile")
    a = ap.parse_args()
    for sh in a.shapes.split(","):
        A, T = (int(v) for v in sh.split(":"))
        for K in (int(v) for v in a.K.split(",")):
            if 4.0 * K * T * A > 12e9:
                continue
            steps = 200 if K <= 100000 else (40 if K <= 500000 else 15)
            row = {"A": A, "T": T, "K": K, "ms": {}}
            for name in a.chains.split(","):
                ms, got, per_step = time_chain(K, T, A, CHAINS[name], steps)
                # a chain the shape does not support falls back silently: tell by the launches
                if name in ("step", "tile") and per_step != 1:
                    continue
                row["ms"][name] = ms
            ms, got, _ = time_chain(K, T, A, capi.FLAG_AUTO_CHAIN, steps)
            row["best"] = min(row["ms"], key=row["ms"].get)
            row["auto"], row["auto_ms"] = got, ms
            row["auto_vs_best"] = ms / row["ms"][row["best"]]
            row["rollout_steps_per_s_best"] = K * T / (row["ms"][row["best"]] * 1e-3)
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
