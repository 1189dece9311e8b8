# Philox-4x32-7 option: tests, then timing of the chains with 7 and 10 rounds at the shard sizes of 1..8 GPUs
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipelined.py tests/test_gpu_step_kernel.py tests/test_bench_contract.py -x -q -m gpu 2>&1 | tail -n 3
python - <<'PY'
import sys, json, numpy as np
sys.path.insert(0, '.')
import mppi_gpu_b200 as m
from mppi_gpu_b200 import capi
cfg = ([1, .5, .75, 0, 0, 0], [1, 1, 1, 5, 5, 5])
for K in (125000, 250000, 500000, 1000000):
    row = {"K": K}
    for name, flags, r in (("auto10", capi.FLAG_AUTO_CHAIN, 10), ("fused10", capi.FLAG_FUSED_SAMPLING, 10),
                           ("fused7", capi.FLAG_FUSED_SAMPLING, 7), ("unfused7", 0, 7), ("auto7", capi.FLAG_AUTO_CHAIN, 7)):
        ctl = m.PointMassModel(K, 200, 0.1, 6, 3, flags=flags, philox_rounds=r)
        ctl.memcpy_set_data(np.zeros(6), np.zeros(600), *cfg)
        for _ in range(5): ctl.get_act()
        ctl.timer_start()
        for _ in range(30): ctl.step_enqueue()
        ms = ctl.timer_stop() / 30
        ctl.step_wait()
        row[name] = round(ms, 4); row[name + "_flags"] = ctl.flags()
        ctl.close()
    print(json.dumps(row), flush=True)
PY
