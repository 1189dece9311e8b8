"""The one-kernel control step (MPPI_FLAG_STEP_KERNEL, csrc/step.cu): rollout warps, TMA
producer and weighted-average consumers on every SM at once, online-softmax merge.

Same bars as tests/test_gpu_parity.py: the noise the GPU drew is dumped through get_inf and
injected into the oracle (the reference's own validation flow); rollout costs, beta and the
argmin index are bit exact, eta / weights / U / next action within 1e-5 relative."""
import numpy as np
import pytest

from conftest import REF_CFG, bits, make_inputs
from test_gpu_parity import _assert_noise_close, _assert_parity, _close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    import mppi_gpu_b200 as m
    return m


def _flags(strict=False):
    from mppi_gpu_b200 import capi
    return capi.FLAG_STEP_KERNEL | (capi.FLAG_STRICT_ARITH if strict else 0)


# shapes: one partial tile; ragged K (pad columns); every A; T*A not a multiple of the 40-row
# TMA box; more tiles than one CTA-round (148 CTAs x 15 warps x 128 samples = 284160)
@pytest.mark.parametrize("strict", [False, True])
@pytest.mark.parametrize("A,K,T,lam", [
    (1, 5, 3, 1.0), (1, 1027, 37, 1.0), (2, 3000, 50, 1.0), (2, 2049, 23, 0.05),
    (3, 5000, 60, 1.0), (3, 129, 7, 2.0), (4, 777, 41, 1.0), (4, 20000, 13, 0.3),
])
def test_step_kernel_matches_oracle(M, oracle, A, K, T, lam, strict):
    cfg = REF_CFG[A]
    x0, U, _ = make_inputs(K, T, A, seed=100 * A + T)
    ctl = M.PointMassModel(K, T, 0.1, 2 * A, A, seed=7 + A, lam=lam, flags=_flags(strict))
    ctl.memcpy_set_data(x0, U, cfg["goal"], cfg["w"])
    p = oracle.make_problem(K, T, A, 0.1, cfg["goal"], cfg["w"], lam=lam,
                            arith=oracle.ARITH_STRICT if strict else oracle.ARITH_FMA)
    for step in range(3):
        pre = ctl.get_u()
        na = ctl.get_act()
        inf = ctl.get_inf()
        info = ctl.step_info()
        assert info["step"] == step + 1
        ref = oracle.step(p, x0, pre, inf["e"])
        _assert_parity(na, inf, info, ref, K, T, A)
        want = oracle.sample_eps(7 + A, step, 0, K, T, A, [0.025] * A)
        _assert_noise_close(inf["e"], want, 0.025)
    assert ctl.launch_count() == 3          # one kernel per control step
    ctl.close()


def _assert_against_float64(oracle, ctl, na, pre, x0, K, T, A, lam, cfg, stride):
    """Large K: the oracle's serial float32 sums (eta, update) are themselves only ~K*2^-24
    accurate, so eta / U are checked against a float64 recomputation from the taps; rollout
    costs stay bit exact against the oracle on every `stride`-th sample."""
    inf = ctl.get_inf(want_e=True)
    info = ctl.step_info()
    cost, e = inf["cost"], inf["e"]
    assert info["argmin"] == int(np.argmin(cost)) and bits(inf["beta"]) == bits(cost.min())
    ex = np.exp(-(cost.astype(np.float64) - float(inf["beta"])) / lam)
    assert abs(ex.sum() - float(inf["nabla"])) <= 1e-5 * ex.sum()
    assert abs(float(inf["weight"].astype(np.float64).sum()) - 1.0) < 1e-4
    num = np.einsum("k,kr->r", ex / ex.sum(), e.reshape(K, T * A).astype(np.float64))
    un = (np.asarray(pre, np.float64).ravel() + num).reshape(T, A)
    assert _close(na, un[0])
    got = inf["u"]
    assert _close(got[:-1], un[1:]) and _close(got[-1], un[-1])
    # the increments themselves (U is O(0.1), the increment O(1e-3)): 1e-4 relative
    inc = np.abs(got[:-1] - un[1:]).max()
    assert inc <= 1e-4 * np.abs(num).max() + 1e-9, (inc, np.abs(num).max())
    sel = np.arange(0, K, stride)
    p = oracle.make_problem(len(sel), T, A, 0.1, cfg["goal"], cfg["w"], lam=lam,
                            arith=oracle.ARITH_FMA)
    S = oracle.rollout_all(p, x0, np.asarray(pre, np.float32), np.ascontiguousarray(e[sel]),
                           nthreads=8)
    assert np.array_equal(bits(S), bits(cost[sel]))


def test_step_kernel_several_rounds(M, oracle):
    """K large enough that every rollout warp runs more than one tile: the consumers fold
    round n while the rollout warps integrate round n+1."""
    K, T, A = 700001, 12, 2
    cfg = REF_CFG[A]
    x0, U, _ = make_inputs(K, T, A, seed=5)
    ctl = M.PointMassModel(K, T, 0.1, 2 * A, A, seed=11, lam=0.5, flags=_flags())
    ctl.memcpy_set_data(x0, U, cfg["goal"], cfg["w"])
    for _ in range(2):
        pre = ctl.get_u()
        na = ctl.get_act()
        _assert_against_float64(oracle, ctl, na, pre, x0, K, T, A, 0.5, cfg, stride=1)
    ctl.close()


@pytest.mark.parametrize("A,K,T", [(3, 40000, 50), (2, 300000, 20)])
def test_step_kernel_equals_kernel_chain(M, A, K, T):
    """Same seed, same inputs: the one-kernel step and the fused two-kernel chain draw the same
    noise and integrate it with the same arithmetic -- costs, beta, argmin identical; U equal to
    rounding (different summation order of the weighted average)."""
    from mppi_gpu_b200 import capi
    cfg = REF_CFG[A]
    x0, U, _ = make_inputs(K, T, A, seed=31)
    out = []
    for fl in (capi.FLAG_FUSED_SAMPLING, capi.FLAG_STEP_KERNEL, 0):
        ctl = M.PointMassModel(K, T, 0.1, 2 * A, A, seed=5, flags=fl)
        ctl.memcpy_set_data(x0, U, cfg["goal"], cfg["w"])
        na = ctl.get_act()
        inf = ctl.get_inf(want_e=True)
        out.append((na, inf, ctl.step_info()))
        ctl.close()
    (na0, i0, s0) = out[0]
    for na, inf, info in out[1:]:
        assert np.array_equal(bits(inf["e"]), bits(i0["e"]))
        assert np.array_equal(bits(inf["cost"]), bits(i0["cost"]))
        assert info["argmin"] == s0["argmin"] and bits(inf["beta"]) == bits(i0["beta"])
        assert _close(inf["nabla"], i0["nabla"], tol=2e-6)
        assert _close(inf["u"], i0["u"], tol=2e-6) and _close(na, na0, tol=2e-6)


def test_step_kernel_is_bitwise_reproducible(M):
    """Static tile -> warp assignment and fixed consumption/merge order: run to run the same bits."""
    K, T, A = 400000, 25, 3
    cfg = REF_CFG[A]
    x0, U, _ = make_inputs(K, T, A, seed=8)
    runs = []
    for _ in range(3):
        ctl = M.PointMassModel(K, T, 0.1, 2 * A, A, seed=2, flags=_flags())
        ctl.memcpy_set_data(x0, U, cfg["goal"], cfg["w"])
        acts = np.array([ctl.get_act() for _ in range(4)])
        runs.append((acts, ctl.get_u(), ctl.step_info()["eta"]))
        ctl.close()
    for acts, u, eta in runs[1:]:
        assert np.array_equal(bits(acts), bits(runs[0][0]))
        assert np.array_equal(bits(u), bits(runs[0][1]))
        assert bits(np.float32(eta)) == bits(np.float32(runs[0][2]))


def test_step_kernel_falls_back_for_injected_noise(M, oracle):
    """Injected noise has no sampling to fuse: the handle runs the kernel chain instead."""
    K, T, A = 3000, 50, 2
    cfg = REF_CFG[A]
    x0, U, eps = make_inputs(K, T, A, seed=3)
    ctl = M.PointMassModel(K, T, 0.1, 2 * A, A, flags=_flags())
    ctl.memcpy_set_data(x0, U, cfg["goal"], cfg["w"])
    ctl.set_noise(eps)
    na = ctl.get_act()
    inf = ctl.get_inf()
    info = ctl.step_info()
    p = oracle.make_problem(K, T, A, 0.1, cfg["goal"], cfg["w"], arith=oracle.ARITH_FMA)
    ref = oracle.step(p, x0, U, eps)
    _assert_parity(na, inf, info, ref, K, T, A)
    assert ctl.launch_count() == 2
    ctl.close()


def test_step_kernel_full_size_point_mass3d(M, oracle):
    """BASELINE.json configs[2] (K=1e6, T=200, A=3) through the one-kernel step: size-independent
    properties from the taps, and the costs of every 997th sample bit exact against the oracle."""
    K, T, A = 1000000, 200, 3
    cfg = REF_CFG[A]
    ctl = M.PointMassModel(K, T, 0.1, 2 * A, A, seed=3, flags=_flags())
    x0 = np.zeros(2 * A, np.float32)
    U = np.zeros((T, A), np.float32)
    ctl.memcpy_set_data(x0, U, cfg["goal"], cfg["w"])
    na = ctl.get_act()
    _assert_against_float64(oracle, ctl, na, U, x0, K, T, A, 1.0, cfg, stride=997)
    ctl.close()


def test_auto_chain_picks_by_work(M):
    """MPPI_FLAG_AUTO_CHAIN: the library resolves the kernel chain from an estimate of what each
    chain costs on the shard (warps on the fullest SM sub-partition, bytes of one eps pass, tiles
    per rollout round; controller.cu: chain_cost) and reports its choice; whatever it picks draws
    the same noise and produces the same costs.  The four shapes are clear cases of the sweep
    profiles/r02_chain_sweep_v2.jsonl."""
    from mppi_gpu_b200 import capi
    A = 2
    cfg = REF_CFG[A]
    want = {(30000, 200): capi.FLAG_PIPELINED_SAMPLING,   # one warp per sub-partition: the unfused chain
            (150000, 200): capi.FLAG_FUSED_SAMPLING,
            (450000, 10): capi.FLAG_FUSED_SAMPLING,       # enough tiles, but a 36 MB step is too short
            (700000, 120): capi.FLAG_STEP_KERNEL}
    for (K, T), chain in want.items():
        ctl = M.PointMassModel(K, T, 0.1, 2 * A, A, seed=4, flags=capi.FLAG_AUTO_CHAIN)
        assert ctl.flags() == chain, (K, T, ctl.flags())
        ctl.memcpy_set_data(np.zeros(4), np.zeros((T, A)), cfg["goal"], cfg["w"])
        na = ctl.get_act()
        cost = ctl.get_inf(want_e=False)["cost"]
        ctl.close()
        ref = M.PointMassModel(K, T, 0.1, 2 * A, A, seed=4, flags=0)
        ref.memcpy_set_data(np.zeros(4), np.zeros((T, A)), cfg["goal"], cfg["w"])
        na0 = ref.get_act()
        assert np.array_equal(bits(ref.get_inf(want_e=False)["cost"]), bits(cost))
        assert _close(na, na0, tol=2e-6)
        ref.close()


def _device_ms(M, K, T, A, flags, steps):
    cfg = REF_CFG[A]
    ctl = M.PointMassModel(K, T, 0.1, 2 * A, A, seed=1, flags=flags)
    ctl.memcpy_set_data(np.zeros(2 * A), np.zeros((T, A)), cfg["goal"], cfg["w"])
    for _ in range(5):
        ctl.get_act()
    best = None
    for _ in range(3):                          # best of three regions: launch noise out
        ctl.timer_start()
        for _ in range(steps):
            ctl.step_enqueue()
        ms = ctl.timer_stop() / steps
        ctl.step_wait()
        best = ms if best is None else min(best, ms)
    got = ctl.flags()
    ctl.close()
    return best, got


@pytest.mark.parametrize("A,T,K", [
    (3, 200, 30000), (3, 200, 125000), (3, 200, 500000), (2, 50, 100000), (2, 50, 500000),
    (1, 200, 250000), (4, 100, 125000), (1, 50, 1000000), (4, 50, 400000),
    # shard sizes between the powers of two (the 6- and 3-GPU shards of K = 1e6): the rollout cost
    # is a step function of the warps on the fullest SM sub-partition
    (3, 200, 166667), (3, 200, 333334), (2, 120, 450000),
])
def test_auto_chain_is_within_5_percent_of_the_best(M, A, T, K):
    """The chain MPPI_FLAG_AUTO_CHAIN resolves to against the three chains timed on the same
    shape (device time of graph-replayed steps).  The comparison is between chain families: the
    latency option the auto chain adds to the unfused chain (pipelined sampling) is masked, it
    trades back-to-back throughput for time after the state arrives."""
    from mppi_gpu_b200 import capi
    steps = 100 if K <= 150000 else 30
    ms = {}
    for name, fl in (("unfused", 0), ("fused", capi.FLAG_FUSED_SAMPLING), ("step", capi.FLAG_STEP_KERNEL)):
        ms[name], _ = _device_ms(M, K, T, A, fl, steps)
    _, chosen = _device_ms(M, K, T, A, capi.FLAG_AUTO_CHAIN, 5)
    family = chosen & ~capi.FLAG_PIPELINED_SAMPLING
    auto_ms, _ = _device_ms(M, K, T, A, family, steps)
    best = min(ms.values())
    assert auto_ms <= 1.05 * best + 1.5e-3, (chosen, auto_ms, ms)


def test_step_kernel_falls_back_when_rows_do_not_fit(M, oracle):
    """T*A too large for the shared-memory row sums: the flag is accepted and the step runs on
    the kernel chain (two kernels with injected-free sampled noise: sample is separate -> 3)."""
    K, T, A = 2000, 700, 4                      # R = 2800 rows: 179 KB of row sums alone
    cfg = REF_CFG[A]
    x0, U, _ = make_inputs(K, T, A, seed=13)
    ctl = M.PointMassModel(K, T, 0.05, 2 * A, A, seed=1, flags=_flags())
    ctl.memcpy_set_data(x0, U, cfg["goal"], cfg["w"])
    na = ctl.get_act()
    inf = ctl.get_inf()
    info = ctl.step_info()
    assert ctl.launch_count() == 3
    p = oracle.make_problem(K, T, A, 0.05, cfg["goal"], cfg["w"], arith=oracle.ARITH_FMA)
    ref = oracle.step(p, x0, U, inf["e"])
    _assert_parity(na, inf, info, ref, K, T, A)
    ctl.close()


@pytest.mark.parametrize("seed", range(12))
def test_random_shapes_every_chain_agrees(M, oracle, seed):
    """Random (K, T, A, lambda, sigma, arithmetic, model) per seed: the three chains of the same
    controller draw the same noise; the oracle on each chain's noise and U confirms costs bit
    for bit and eta / weights / U within 1e-5."""
    from mppi_gpu_b200 import capi
    rs = np.random.RandomState(1000 + seed)
    A = int(rs.randint(1, 5))
    T = int(rs.randint(1, 90))
    K = int(rs.choice([rs.randint(1, 300), rs.randint(300, 5000), rs.randint(5000, 60000)]))
    lam = float(rs.choice([0.05, 0.5, 1.0, 7.0]))
    sigma = float(rs.choice([0.01, 0.025, 0.3]))
    strict = bool(rs.randint(0, 2))
    gains = None
    if rs.randint(0, 2):
        gains = (np.array([1.0, 0.1, rs.uniform(-0.05, 0.05), rs.uniform(0.85, 1.0)], np.float32),
                 np.array([rs.uniform(0.0, 0.02), rs.uniform(0.05, 0.3)], np.float32))
    cfg = REF_CFG[A]
    x0, U, _ = make_inputs(K, T, A, seed=seed)
    base = capi.FLAG_STRICT_ARITH if strict else 0
    kw = dict(seed=seed, lam=lam, sigma=sigma)
    if gains is not None:
        kw.update(state_gain=gains[0], act_gain=gains[1])
    p = oracle.make_problem(K, T, A, 0.1, cfg["goal"], cfg["w"], lam=lam, gains=gains,
                            arith=oracle.ARITH_STRICT if strict else oracle.ARITH_FMA)
    first = None
    for chain in (capi.FLAG_STEP_KERNEL, capi.FLAG_FUSED_SAMPLING, 0):
        ctl = M.PointMassModel(K, T, 0.1, 2 * A, A, flags=base | chain, **kw)
        ctl.memcpy_set_data(x0, U, cfg["goal"], cfg["w"])
        ctl.get_act()                                   # step 0
        pre = ctl.get_u()
        na = ctl.get_act()                              # step 1, from the updated U
        inf, info = ctl.get_inf(), ctl.step_info()
        ctl.close()
        # U of step 0 agrees between chains to rounding only, so every chain's step 1 is checked
        # through the oracle run from that chain's own U
        ref = oracle.step(p, x0, pre, inf["e"])
        _assert_parity(na, inf, info, ref, K, T, A)
        if first is None:
            first = (pre, inf["e"])
        else:
            assert np.array_equal(bits(inf["e"]), bits(first[1]))
            assert _close(pre, first[0], tol=2e-6)
