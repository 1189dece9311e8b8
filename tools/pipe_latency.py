"""Closed-loop latency of get_act with a plant that takes `delay` microseconds between control
steps (development aid for MPPI_FLAG_PIPELINED_SAMPLING: the sampler of the next step runs
during the plant's turn, off the critical path)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import mppi_gpu_b200 as m

K, T, A = int(sys.argv[1]) if len(sys.argv) > 1 else 100000, 200, int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = {2: ([1, 0, 0, 0], [1, 1, 50, 50]), 3: ([1, .5, .75, 0, 0, 0], [1, 1, 1, 5, 5, 5])}[A]
for flags in (0, m.capi.FLAG_PIPELINED_SAMPLING):
    ctl = m.PointMassModel(K, T, 0.1, 2 * A, A, flags=flags)
    ctl.memcpy_set_data(np.zeros(2 * A), np.zeros(T * A), *cfg)
    x = np.zeros(2 * A, np.float32)
    out = np.zeros(A, np.float32)
    for delay_us in (0, 20, 40, 80, 200):
        lat = []
        for i in range(600):
            ctl.set_x(x)
            t0 = time.perf_counter()
            ctl.get_act(out)
            t1 = time.perf_counter()
            lat.append(t1 - t0)
            while time.perf_counter() - t1 < delay_us * 1e-6:     # the plant's turn
                pass
        lat = np.array(lat[100:]) * 1e6
        print(json.dumps({"K": K, "A": A, "flags": flags, "plant_us": delay_us,
                          "p50_us": round(float(np.percentile(lat, 50)), 1),
                          "p99_us": round(float(np.percentile(lat, 99)), 1)}))
    ctl.close()
