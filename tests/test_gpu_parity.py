"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the
C ABI (libmppi_b200.so) via mppi_gpu_b200.PointMassModel and is checked against the
CPU oracle on identical injected noise.

Bars (BASELINE.json north_star):
  * rollout costs S_k: bit-exact against the oracle in BOTH arithmetic modes
    (strict == reference host build, fma == reference device build contraction);
  * beta bit-exact, argmin index exact;
  * eta, weights, updated U, next action: |diff| <= 1e-5 * max(1, |value|)
    (FP32, reordered reductions);
  * sampled noise: Philox integers are exact by construction; the Box-Muller floats use
    MUFU approximations (lg2/sqrt/sin/cos): |eps_gpu - eps_oracle| <= 2e-5*sigma for all but
    <= 1e-5 of the draws, and <= 1e-3*sigma always (lg2.approx has 2^-22 ABSOLUTE error on
    (0.5,2), so the radius sqrt(-2 ln u) is only that accurate when u -> 1, i.e. r -> 0).
"""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, REF_CFG, bits, golden_gains, load_golden, make_inputs

pytestmark = pytest.mark.gpu

RTOL_U = 1e-5


def _close(a, b, tol=RTOL_U):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    assert a.shape == b.shape, (a.shape, b.shape)
    return np.all(np.abs(a - b) <= tol * np.maximum(1.0, np.abs(b)))


def _assert_noise_close(e, want, sigma):
    d = np.abs(e - want) / sigma
    assert d.max() < 1e-3, d.max()
    # (at most 2 draws for small arrays, where one draw already exceeds 1e-5 of them)
    assert np.mean(d > 2e-5) <= max(1e-5, 2.0 / d.size), np.mean(d > 2e-5)


@pytest.fixture(scope="module")
def M():
    import mppi_gpu_b200 as m
    return m


def _run_case(M, oracle, K, T, A, dt, goal, w, x0, U, eps, lam=1.0, strict=False, flags=0,
              inv_sigma=1.0):
    from mppi_gpu_b200 import capi
    fl = flags | (capi.FLAG_STRICT_ARITH if strict else 0)
    ctl = M.PointMassModel(K, T, dt, 2 * A, A, lam=lam, flags=fl, inv_sigma=inv_sigma)
    ctl.memcpy_set_data(x0, U, goal, w)
    ctl.set_noise(eps)
    next_act = ctl.get_act()
    inf = ctl.get_inf(want_x=(K * T <= 200000))
    info = ctl.step_info()
    p = oracle.make_problem(K, T, A, dt, goal, w, lam=lam,
                            inv_s=np.broadcast_to(np.float32(inv_sigma), (A,)),
                            arith=oracle.ARITH_STRICT if strict else oracle.ARITH_FMA)
    ref = oracle.step(p, x0, U, eps)
    ctl.close()
    return next_act, inf, info, ref, p


def _assert_parity(next_act, inf, info, ref, K, T, A):
    # (2) rollout costs: bit exact
    assert np.array_equal(bits(inf["cost"]), bits(ref["S"])), \
        f"max |dS| = {np.abs(inf['cost'] - ref['S']).max()}"
    # (3) beta / argmin exact
    assert bits(inf["beta"]) == bits(ref["beta"])
    assert info["argmin"] == ref["argmin"]
    # eta, weights within tolerance
    assert _close(inf["nabla"], ref["eta"])
    assert _close(inf["weight"], ref["weights"])
    assert abs(float(inf["weight"].astype(np.float64).sum()) - 1.0) < 1e-4
    # (4) U (post shift) and next action
    assert _close(inf["u"].ravel(), ref["U"]), \
        f"max |dU| = {np.abs(inf['u'].ravel() - ref['U'].ravel()).max()}"
    assert _close(next_act, ref["next_act"])
    # noise tap returns exactly what was injected, in the reference layout
    assert inf["e"].shape == (K, T, A)


@pytest.mark.parametrize("strict", [False, True])
@pytest.mark.parametrize("A", [1, 2, 3, 4])
@pytest.mark.parametrize("K,T", [(3000, 50), (3, 12), (59, 99), (1, 1), (5, 3), (257, 37),
                                 (1023, 20), (1025, 7), (4096, 200)])
def test_step_matches_oracle(M, oracle, K, T, A, strict):
    cfg = REF_CFG[A]
    x0, U, eps = make_inputs(K, T, A, seed=K * 31 + T * 7 + A)
    next_act, inf, info, ref, p = _run_case(M, oracle, K, T, A, 0.1, cfg["goal"], cfg["w"], x0, U,
                                            eps, strict=strict)
    _assert_parity(next_act, inf, info, ref, K, T, A)
    assert np.array_equal(inf["e"], eps)
    if inf["x"] is not None:
        _, xt = oracle.rollout_all(p, x0, U, eps, want_traj=True)
        assert np.array_equal(inf["x"], xt)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_vectors(M, oracle, name):
    """Committed vectors produced by the reference's own sources (tests/golden/make_golden.py)."""
    g = load_golden(name)
    K, T, A = int(g["K"]), int(g["T"]), int(g["A"])
    for strict in (True, False):
        gains = golden_gains(g) or (None, None)
        ctl = M.PointMassModel(K, T, float(g["dt"]), 2 * A, A, flags=1 if strict else 0,
                               state_gain=gains[0], act_gain=gains[1])
        ctl.memcpy_set_data(g["x0"], g["U"], g["goal"], g["w"])
        ctl.set_noise(g["eps"])
        na = ctl.get_act()
        inf = ctl.get_inf(want_x=("x_ref" in g and strict))
        info = ctl.step_info()
        ctl.close()
        sfx = "" if strict else "_fma"
        assert np.array_equal(bits(inf["cost"]), bits(g["S_ref"] if strict else g["S_fma"]))
        assert info["argmin"] == int(g["argmin" + sfx])
        assert bits(inf["beta"]) == bits(g["beta" + sfx])
        assert _close(inf["nabla"], g["eta" + sfx])
        assert _close(inf["u"].ravel(), g["U_next" + sfx])
        assert _close(na, g["next_act" + sfx])
        if strict:
            assert _close(inf["weight"], g["weights"])
            if "x_ref" in g:
                assert np.array_equal(inf["x"], g["x_ref"])


@pytest.mark.parametrize("lam", [1.5, 10.0, 100.0])
def test_temperature_exercises_the_average(M, oracle, lam):
    """Large lambda flattens the weights so that every sample contributes to U."""
    K, T, A = 2048, 50, 2
    cfg = REF_CFG[A]
    x0, U, eps = make_inputs(K, T, A, seed=99, sigma=0.25)
    next_act, inf, info, ref, _ = _run_case(M, oracle, K, T, A, 0.1, cfg["goal"], cfg["w"], x0, U,
                                            eps, lam=lam)
    _assert_parity(next_act, inf, info, ref, K, T, A)
    if lam >= 10:
        assert inf["nabla"] > 5          # many effective samples


def test_inv_sigma_reaches_control_cost(M, oracle):
    K, T, A = 512, 30, 2
    cfg = REF_CFG[A]
    x0, U, eps = make_inputs(K, T, A, seed=4, u_scale=1.0)
    next_act, inf, info, ref, _ = _run_case(M, oracle, K, T, A, 0.1, cfg["goal"], cfg["w"], x0, U,
                                            eps, lam=2.0, inv_sigma=16.0, strict=True)
    _assert_parity(next_act, inf, info, ref, K, T, A)


def test_reference_update_act_fixture(M, oracle):
    """The reference's own unit-test fixture for the averaging pass (src/test.cu:77-105):
    e_i = 0.25 i, U_i = 0.75 i, A = 2 -- driven through the full step with equal costs."""
    K, T, A = 59, 99, 2
    idx = np.arange(K * T * A, dtype=np.float64)
    eps = (0.25 * idx * 1e-4).astype(np.float32).reshape(K, T, A)   # scaled to stay finite
    U = (0.75 * np.arange(T * A) * 1e-3).astype(np.float32)
    x0 = np.zeros(4, np.float32)
    goal = np.zeros(4, np.float32)
    w = np.zeros(4, np.float32)                                      # S_k = control cost only
    next_act, inf, info, ref, _ = _run_case(M, oracle, K, T, A, 0.1, goal, w, x0, U, eps,
                                            lam=1e6, strict=True)
    _assert_parity(next_act, inf, info, ref, K, T, A)


def _plant(x, act, g, b, A):
    """ideal double integrator (the model of src/model_missmatch.cpp:26-38)"""
    pos, vel = x[:A].astype(np.float32), x[A:].astype(np.float32)
    act = np.asarray(act, np.float32)
    return np.concatenate([pos + g[1] * vel + b[0] * act, vel + b[1] * act]).astype(np.float32)


def test_closed_loop_200_steps_tracks_oracle(M, oracle):
    """Receding-horizon loop of src/main.cu:326-371 (get_u, get_act, plant, set_x) for 200
    control steps on identical injected noise.
      * per step, from the GPU's own (x, U): next action / U' within 1e-5, argmin exact;
      * free-running: the GPU loop and an independent oracle loop (own plant copy, own U)
        stay within CLOSED_LOOP_TOL = 1e-3 * max(1, |x|) in state over all 200 steps."""
    CLOSED_LOOP_TOL = 1e-3
    K, T, A = 512, 30, 2
    cfg = REF_CFG[A]
    rs = np.random.RandomState(5)
    ctl = M.PointMassModel(K, T, 0.1, 2 * A, A, flags=1)
    x_gpu = np.zeros(4, np.float32)
    x_orc = x_gpu.copy()
    U0 = np.zeros((T, A), np.float32)
    ctl.memcpy_set_data(x_gpu, U0, cfg["goal"], cfg["w"])
    p = oracle.make_problem(K, T, A, 0.1, cfg["goal"], cfg["w"])
    U_orc = U0.copy()
    g, b = oracle.gains(0.1)
    worst = 0.0
    for step in range(200):
        eps = (0.25 * rs.standard_normal((K, T, A))).astype(np.float32)
        ctl.set_noise(eps)
        pre = ctl.get_u()                              # main.cu:327
        na = ctl.get_act()                             # main.cu:330
        info = ctl.step_info()
        one = oracle.step(p, x_gpu, pre, eps)          # same state, same U, same noise
        assert _close(na, one["next_act"])
        assert info["argmin"] == one["argmin"] and info["step"] == step + 1
        assert _close(ctl.get_u(), one["U"])
        free = oracle.step(p, x_orc, U_orc, eps)       # independent oracle loop
        U_orc = free["U"]
        x_gpu = _plant(x_gpu, na, g, b, A)
        x_orc = _plant(x_orc, free["next_act"], g, b, A)
        err = np.max(np.abs(x_gpu - x_orc) / np.maximum(1.0, np.abs(x_orc)))
        worst = max(worst, float(err))
        assert err <= CLOSED_LOOP_TOL, (step, err)
        ctl.set_x(x_gpu)                               # main.cu:371
    # the controller actually drove the mass towards the goal position (1, 0)
    assert x_gpu[0] > 0.4 and abs(x_gpu[1]) < 0.2, x_gpu
    print("closed loop worst relative state deviation:", worst)
    ctl.close()


def test_reinit_and_clamp_flags(M, oracle):
    from mppi_gpu_b200 import capi
    K, T, A = 512, 20, 2
    cfg = REF_CFG[A]
    x0, U, eps = make_inputs(K, T, A, seed=8, sigma=0.5, u_scale=0.5)
    ctl = M.PointMassModel(K, T, 0.1, 2 * A, A, flags=capi.FLAG_CLAMP_ACTIONS |
                           capi.FLAG_REINIT_INIT_ACT | capi.FLAG_STRICT_ARITH,
                           init_act=[0.1, 0.2], max_act=[0.3, 0.4], lam=50.0)
    ctl.memcpy_set_data(x0, U, cfg["goal"], cfg["w"])
    ctl.set_noise(eps)
    na = ctl.get_act()
    u = ctl.get_u()
    p = oracle.make_problem(K, T, A, 0.1, cfg["goal"], cfg["w"], lam=50.0)
    S = oracle.rollout_all(p, x0, U, eps)
    b, _ = oracle.beta(S)
    ex = oracle.exp(S, 50.0, b)
    w = oracle.weights(S, 50.0, b, oracle.eta(ex)[0])
    un = oracle.update_act(U, w, eps, K, T, A).reshape(T, A)
    un = np.clip(un, [-0.3, -0.4], [0.3, 0.4])
    assert _close(na, un[0])
    assert _close(u[:-1], un[1:])
    assert np.allclose(u[-1], [0.1, 0.2])
    ctl.close()


# ------------------------------------------------------------------ (1) sampling
@pytest.mark.parametrize("rounds", [10, 7])
@pytest.mark.parametrize("A", [1, 2, 3])
def test_sampler_matches_oracle_stream(M, oracle, A, rounds):
    K, T = 1027, 33
    sig = [0.025, 0.1, 0.3][:A]
    ctl = M.PointMassModel(K, T, 0.1, 2 * A, A, sigma=sig, seed=0x1234567890ABCDEF, philox_rounds=rounds)
    cfg = REF_CFG[A]
    ctl.memcpy_set_data(np.zeros(2 * A), np.zeros(T * A), cfg["goal"], cfg["w"])
    for step in (0, 5, 2 ** 33 + 1):
        ctl.sample_only(step)
        e = ctl.get_inf()["e"]
        want = oracle.sample_eps(0x1234567890ABCDEF, step, 0, K, T, A, sig, rounds=rounds)
        s = np.asarray(sig, np.float32)[None, None, :]
        _assert_noise_close(e, want, s)
    ctl.close()


def test_sampler_statistics(M):
    K, T, A = 100000, 20, 2
    ctl = M.PointMassModel(K, T, 0.1, 4, 2, sigma=[0.025, 0.05], seed=1)
    ctl.memcpy_set_data(np.zeros(4), np.zeros(T * A), REF_CFG[2]["goal"], REF_CFG[2]["w"])
    ctl.sample_only(0)
    e = ctl.get_inf()["e"].astype(np.float64)
    for a, s in enumerate((0.025, 0.05)):
        col = e[:, :, a].ravel() / s
        assert abs(col.mean()) < 4 / np.sqrt(col.size)
        assert abs(col.std() - 1) < 5e-3
        assert abs((col ** 3).mean()) < 0.02 and abs((col ** 4).mean() - 3) < 0.05
    # consecutive steps are independent streams
    ctl.sample_only(1)
    e1 = ctl.get_inf()["e"].astype(np.float64)
    c = np.corrcoef(e.ravel(), e1.ravel())[0, 1]
    assert abs(c) < 5 / np.sqrt(e.size)
    ctl.close()


@pytest.mark.parametrize("fused", [False, True])
def test_sampled_step_is_self_consistent(M, oracle, fused):
    """Full sampled step (Philox path, optionally fused into the rollout): dump the noise the
    GPU drew, inject it into the oracle -- the reference's own validation flow."""
    from mppi_gpu_b200 import capi
    K, T, A = 5000, 60, 3
    cfg = REF_CFG[A]
    x0, U, _ = make_inputs(K, T, A, seed=21)
    ctl = M.PointMassModel(K, T, 0.1, 2 * A, A, seed=42,
                           flags=capi.FLAG_FUSED_SAMPLING if fused else 0)
    ctl.memcpy_set_data(x0, U, cfg["goal"], cfg["w"])
    p = oracle.make_problem(K, T, A, 0.1, cfg["goal"], cfg["w"], arith=oracle.ARITH_FMA)
    for step in range(3):
        pre = ctl.get_u()                       # the U this step starts from (GPU's own)
        na = ctl.get_act()
        inf = ctl.get_inf()
        info = ctl.step_info()
        ref = oracle.step(p, x0, pre, inf["e"])
        _assert_parity(na, inf, info, ref, K, T, A)
        want = oracle.sample_eps(42, step, 0, K, T, A, [0.025] * A)
        _assert_noise_close(inf["e"], want, 0.025)
    ctl.close()
    # fused and unfused sampling draw the same noise: checked through the oracle stream above


def test_graph_and_direct_launch_agree_bitwise(M):
    from mppi_gpu_b200 import capi
    K, T, A = 3000, 50, 2
    cfg = REF_CFG[A]
    x0, U, _ = make_inputs(K, T, A, seed=2)
    res = []
    for fl in (0, capi.FLAG_NO_GRAPH):
        ctl = M.PointMassModel(K, T, 0.1, 4, 2, seed=9, flags=fl)
        ctl.memcpy_set_data(x0, U, cfg["goal"], cfg["w"])
        acts = [ctl.get_act() for _ in range(5)]
        res.append((np.array(acts), ctl.get_u(), ctl.get_inf()["cost"]))
        ctl.close()
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(bits(a), bits(b))


def test_profiling_mode_reports_every_kernel(M):
    from mppi_gpu_b200 import capi
    K, T, A = 20000, 50, 2
    cfg = REF_CFG[A]
    x0, U, _ = make_inputs(K, T, A, seed=6)
    results = {}
    for flags, names, per_step in (
            (capi.FLAG_SPLIT_KERNELS, ("sample", "rollout", "weights", "average", "finalize"), 5),
            (0, ("sample", "rollout", "average"), 3),
            (capi.FLAG_FUSED_SAMPLING, ("rollout", "average"), 2)):
        ctl = M.PointMassModel(K, T, 0.1, 4, 2, flags=flags, seed=3)
        ctl.memcpy_set_data(x0, U, cfg["goal"], cfg["w"])
        ctl.get_act()
        results[flags] = (ctl.get_u(), ctl.get_inf(want_e=False)["cost"], ctl.step_info())
        ctl.set_profiling(True)
        for _ in range(3):
            ctl.get_act()
        kt = ctl.kernel_times()
        ctl.set_profiling(False)
        for name, (ms, n) in kt.items():
            if name in names:
                assert n == 3 and ms > 0, (flags, name)
            else:
                assert n == 0, (flags, name)
        assert ctl.launch_count() == 4 * per_step
        ctl.close()
    # merged, split and fused chains are the same computation (compared after the first
    # step, from the same U): identical costs, beta and argmin; U agrees to rounding (eta partials are grouped differently before they enter
    # the fixed-point accumulator)
    ref = results[0]
    for flags, (u, cost, info) in results.items():
        assert np.array_equal(bits(cost), bits(ref[1])), flags
        assert info["argmin"] == ref[2]["argmin"] and info["step"] == ref[2]["step"], flags
        assert bits(info["beta"]) == bits(ref[2]["beta"]), flags
        assert _close(u, ref[0], tol=2e-6), flags


# ------------------------------------------------------------------ full size (BASELINE configs)
def test_full_size_point_mass3d_properties(M, oracle):
    """point_mass3d, K=1e6, T=200 (BASELINE.json configs[2]) on one GPU: size-independent
    properties + a full oracle comparison of costs on a strided subset of samples."""
    K, T, A = 1000000, 200, 3
    cfg = REF_CFG[A]
    ctl = M.PointMassModel(K, T, 0.1, 2 * A, A, seed=3)
    x0 = np.zeros(2 * A, np.float32)
    U = np.zeros((T, A), np.float32)
    ctl.memcpy_set_data(x0, U, cfg["goal"], cfg["w"])
    na = ctl.get_act()
    inf = ctl.get_inf(want_e=True)
    info = ctl.step_info()
    cost, e, wgt = inf["cost"], inf["e"], inf["weight"]
    # beta / argmin are the exact min / first argmin of the cost array
    assert info["argmin"] == int(np.argmin(cost)) and bits(inf["beta"]) == bits(cost.min())
    # eta = sum exp(-(S-beta)), weights sum to one
    ex = np.exp(-(cost.astype(np.float64) - float(inf["beta"])))
    assert abs(ex.sum() - float(inf["nabla"])) <= 1e-5 * ex.sum()
    assert abs(wgt.astype(np.float64).sum() - 1) < 1e-4
    # U' = shift(U + sum_k w_k eps_k) against a float64 recomputation from the taps
    num = np.einsum("k,kr->r", ex / ex.sum(), e.reshape(K, T * A).astype(np.float64))
    un = (U.ravel() + num).reshape(T, A)
    assert _close(na, un[0])
    got = inf["u"]
    assert _close(got[:-1], un[1:]) and _close(got[-1], un[-1])
    # rollout costs of every 997th sample, bit exact against the oracle
    sel = np.arange(0, K, 997)
    p = oracle.make_problem(len(sel), T, A, 0.1, cfg["goal"], cfg["w"], arith=oracle.ARITH_FMA)
    S = oracle.rollout_all(p, x0, U, np.ascontiguousarray(e[sel]), nthreads=8)
    assert np.array_equal(bits(S), bits(cost[sel]))
    # sampled noise moments at scale
    z = e[:: 50].astype(np.float64).ravel() / 0.025
    assert abs(z.mean()) < 5 / np.sqrt(z.size) and abs(z.std() - 1) < 2e-3
    ctl.close()


def test_seven_round_philox_chains(M, oracle):
    """mppi_params.philox_rounds = 7 (Random123's philox4x32_R<7>): the fused and the unfused chain
    draw the same noise, bit for bit, and it is the oracle's 7-round stream; costs / argmin / U as
    always; a one-kernel request (the one-kernel steps exist with ten rounds only) runs the fused chain;
    ten rounds give a different stream."""
    from mppi_gpu_b200 import capi
    K, T, A = 5000, 60, 3
    cfg = REF_CFG[A]
    x0, U, _ = make_inputs(K, T, A, seed=29)
    p = oracle.make_problem(K, T, A, 0.1, cfg["goal"], cfg["w"], arith=oracle.ARITH_FMA)
    noise = {}
    for name, flags in (("unfused", 0), ("fused", capi.FLAG_FUSED_SAMPLING), ("step", capi.FLAG_STEP_KERNEL)):
        ctl = M.PointMassModel(K, T, 0.1, 2 * A, A, seed=42, flags=flags, philox_rounds=7)
        if name == "step":
            assert ctl.flags() & capi.FLAG_FUSED_SAMPLING and not ctl.flags() & capi.FLAG_STEP_KERNEL
        ctl.memcpy_set_data(x0, U, cfg["goal"], cfg["w"])
        for step in range(2):
            pre = ctl.get_u()
            na = ctl.get_act()
            inf = ctl.get_inf()
            info = ctl.step_info()
            ref = oracle.step(p, x0, pre, inf["e"])
            _assert_parity(na, inf, info, ref, K, T, A)
            want = oracle.sample_eps(42, step, 0, K, T, A, [0.025] * A, rounds=7)
            _assert_noise_close(inf["e"], want, 0.025)
        noise[name] = inf["e"].copy()
        ctl.close()
    assert np.array_equal(bits(noise["unfused"]), bits(noise["fused"]))
    assert np.array_equal(bits(noise["step"]), bits(noise["fused"]))
    ten = M.PointMassModel(K, T, 0.1, 2 * A, A, seed=42, flags=capi.FLAG_FUSED_SAMPLING)
    ten.memcpy_set_data(x0, U, cfg["goal"], cfg["w"])
    ten.get_act(); ten.get_act()
    assert not np.array_equal(bits(ten.get_inf()["e"]), bits(noise["fused"]))
    ten.close()
    with pytest.raises(Exception):
        M.PointMassModel(K, T, 0.1, 2 * A, A, philox_rounds=8)
