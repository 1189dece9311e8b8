import sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "tests")))
import numpy as np
import mppi_gpu_b200 as m
from mppi_gpu_b200 import capi
from conftest import REF_CFG, make_inputs
K, T, A = 700001, 12, 2
lam = 0.5
cfg = REF_CFG[A]
x0, U, _ = make_inputs(K, T, A, seed=5)
for flags in (capi.FLAG_STEP_KERNEL, capi.FLAG_FUSED_SAMPLING):
    ctl = m.PointMassModel(K, T, 0.1, 2 * A, A, seed=11, lam=lam, flags=flags)
    ctl.memcpy_set_data(x0, U, cfg["goal"], cfg["w"])
    pre = ctl.get_u()
    na = ctl.get_act()
    inf = ctl.get_inf()
    info = ctl.step_info()
    S = inf["cost"].astype(np.float64); e = inf["e"].reshape(K, T * A).astype(np.float64)
    beta = S.min()
    w = np.exp(-(S - beta) / lam)
    eta = w.sum()
    print("flags", flags, "eta gpu", info["eta"], "eta f64", eta, "rel", (info["eta"] - eta) / eta)
    num = w @ e            # [R]
    # gpu increment: U_new(before shift) - pre ; u after shift: u[t] = Unew[t+1]
    got = inf["u"].astype(np.float64)
    inc_gpu = np.empty((T, A)); inc_gpu[1:] = got[:-1] - pre[1:]; inc_gpu[0] = na - pre[0]
    inc_ref = (num / eta).reshape(T, A)
    err = (inc_gpu - inc_ref).ravel()
    print(" max |inc err|", np.abs(err).max(), "max |inc|", np.abs(inc_ref).max())
    # which 128-sample tile explains err * eta?
    ntile = (K + 127) // 128
    pad = ntile * 128 - K
    we = (w[:, None] * e)
    we = np.concatenate([we, np.zeros((pad, T * A))]).reshape(ntile, 128, T * A).sum(1)   # [ntile][R]
    target = err * eta
    alpha = (we @ target) / np.maximum((we * we).sum(1), 1e-300)
    res = np.linalg.norm(target[None, :] - alpha[:, None] * we, axis=1)
    best = np.argsort(res)[:5]
    print(" |target|", np.linalg.norm(target), "best tiles", best, "alpha", alpha[best], "res", res[best])
    wt = np.concatenate([w, np.zeros(pad)]).reshape(ntile, 128).sum(1)
    print(" eta deficit", eta - float(info["eta"]), "tile weight sums of best", wt[best])
    ctl.close()
