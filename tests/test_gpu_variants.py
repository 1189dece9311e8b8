"""Every kernel variant that the size heuristics would otherwise only pick for some shapes,
forced through environment overrides and checked against the oracle (compute-sanitizer is
closed on this GPU pool, so out-of-bounds behaviour is probed with ragged shapes, every
variant, and bit-exact comparisons instead)."""
import os

import numpy as np
import pytest

from conftest import REF_CFG, bits, make_inputs

pytestmark = pytest.mark.gpu

VARIANTS = [
    ("ldg_spt1", {"MPPI_ROLLOUT_TMA": "0", "MPPI_ROLLOUT_SPT": "1"}, 0),
    ("ldg_spt2_packed", {"MPPI_ROLLOUT_TMA": "0", "MPPI_ROLLOUT_SPT": "2"}, 0),
    ("ldg_spt4_packed", {"MPPI_ROLLOUT_TMA": "0", "MPPI_ROLLOUT_SPT": "4"}, 0),
    ("tma_w64", {"MPPI_ROLLOUT_TMA": "1", "MPPI_ROLLOUT_TMA_W": "64"}, 0),
    ("tma_w128", {"MPPI_ROLLOUT_TMA": "1", "MPPI_ROLLOUT_TMA_W": "128"}, 0),
    ("tma_w256", {"MPPI_ROLLOUT_TMA": "1", "MPPI_ROLLOUT_TMA_W": "256"}, 0),
    ("split_kernels", {}, 64),
    ("no_graph", {}, 16),
]


@pytest.mark.parametrize("strict", [False, True])
@pytest.mark.parametrize("name,env,flags", VARIANTS, ids=[v[0] for v in VARIANTS])
@pytest.mark.parametrize("A,K,T", [(1, 1027, 37), (2, 2049, 23), (3, 3000, 50), (4, 777, 41), (3, 5, 3)])
def test_variant_matches_oracle(oracle, name, env, flags, A, K, T, strict):
    import mppi_gpu_b200 as m
    cfg = REF_CFG[A]
    x0, U, eps = make_inputs(K, T, A, seed=A * 1000 + K, sigma=0.2)
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        ctl = m.PointMassModel(K, T, 0.1, 2 * A, A, flags=flags | (1 if strict else 0), lam=3.0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    ctl.memcpy_set_data(x0, U, cfg["goal"], cfg["w"])
    ctl.set_noise(eps)
    na = ctl.get_act()
    inf = ctl.get_inf(want_e=False)
    info = ctl.step_info()
    ctl.close()
    p = oracle.make_problem(K, T, A, 0.1, cfg["goal"], cfg["w"], lam=3.0,
                            arith=oracle.ARITH_STRICT if strict else oracle.ARITH_FMA)
    ref = oracle.step(p, x0, U, eps)
    assert np.array_equal(bits(inf["cost"]), bits(ref["S"])), np.abs(inf["cost"] - ref["S"]).max()
    assert info["argmin"] == ref["argmin"] and bits(inf["beta"]) == bits(ref["beta"])
    assert np.allclose(inf["u"].ravel(), ref["U"].ravel(), rtol=1e-5, atol=1e-6)
    assert np.allclose(na, ref["next_act"], rtol=1e-5, atol=1e-6)


GAINS = (np.array([0.97, 0.12, 0.03, 0.9], np.float32), np.array([0.02, 0.3], np.float32))


@pytest.mark.parametrize("strict", [False, True])
@pytest.mark.parametrize("name,env,flags", VARIANTS[:6], ids=[v[0] for v in VARIANTS[:6]])
@pytest.mark.parametrize("A,K,T", [(1, 1027, 37), (2, 2049, 23), (3, 3000, 50), (4, 777, 41)])
def test_linear_axis_model_matches_oracle(oracle, name, env, flags, A, K, T, strict):
    """The second dynamics functor (MPPI_MODEL_LINEAR_AXIS, caller-given gains) through every
    rollout kernel variant: costs bit exact against the oracle, which tests/test_oracle.py pins
    to the reference's own step() with the same gains."""
    import mppi_gpu_b200 as m
    cfg = REF_CFG[A]
    x0, U, eps = make_inputs(K, T, A, seed=A * 77 + K, sigma=0.2)
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        ctl = m.PointMassModel(K, T, 0.1, 2 * A, A, flags=flags | (1 if strict else 0), lam=3.0,
                               state_gain=GAINS[0], act_gain=GAINS[1])
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    ctl.memcpy_set_data(x0, U, cfg["goal"], cfg["w"])
    ctl.set_noise(eps)
    na = ctl.get_act()
    inf = ctl.get_inf(want_e=False, want_x=(name == "ldg_spt1"))
    info = ctl.step_info()
    ctl.close()
    p = oracle.make_problem(K, T, A, 0.1, cfg["goal"], cfg["w"], lam=3.0, gains=GAINS,
                            arith=oracle.ARITH_STRICT if strict else oracle.ARITH_FMA)
    ref = oracle.step(p, x0, U, eps)
    assert np.array_equal(bits(inf["cost"]), bits(ref["S"])), np.abs(inf["cost"] - ref["S"]).max()
    assert info["argmin"] == ref["argmin"] and bits(inf["beta"]) == bits(ref["beta"])
    assert np.allclose(inf["u"].ravel(), ref["U"].ravel(), rtol=1e-5, atol=1e-6)
    assert np.allclose(na, ref["next_act"], rtol=1e-5, atol=1e-6)
    if inf["x"] is not None:
        _, xt = oracle.rollout_all(p, x0, U, eps, want_traj=True)
        assert np.array_equal(inf["x"], xt)


@pytest.mark.parametrize("flags", [32, 128], ids=["fused", "step_kernel"])
def test_linear_axis_model_sampled_chains(oracle, flags):
    """Sampled noise through the fused chain and the one-kernel step with the LinearAxis model."""
    import mppi_gpu_b200 as m
    K, T, A = 6000, 40, 3
    cfg = REF_CFG[A]
    x0, U, _ = make_inputs(K, T, A, seed=19)
    ctl = m.PointMassModel(K, T, 0.1, 2 * A, A, seed=9, flags=flags, state_gain=GAINS[0],
                           act_gain=GAINS[1])
    ctl.memcpy_set_data(x0, U, cfg["goal"], cfg["w"])
    p = oracle.make_problem(K, T, A, 0.1, cfg["goal"], cfg["w"], gains=GAINS, arith=oracle.ARITH_FMA)
    for _ in range(2):
        pre = ctl.get_u()
        na = ctl.get_act()
        inf = ctl.get_inf()
        info = ctl.step_info()
        ref = oracle.step(p, x0, pre, inf["e"])
        assert np.array_equal(bits(inf["cost"]), bits(ref["S"]))
        assert info["argmin"] == ref["argmin"] and bits(inf["beta"]) == bits(ref["beta"])
        assert np.allclose(inf["u"].ravel(), ref["U"].ravel(), rtol=1e-5, atol=1e-6)
        assert np.allclose(na, ref["next_act"], rtol=1e-5, atol=1e-6)
    ctl.close()



@pytest.mark.parametrize("strict", [False, True])
@pytest.mark.parametrize("name,env,flags", VARIANTS[:6], ids=[v[0] for v in VARIANTS[:6]])
@pytest.mark.parametrize("A,K,T", [(1, 1027, 37), (2, 2049, 23), (3, 3000, 50), (4, 777, 41)])
def test_terminal_weights_match_oracle(oracle, name, env, flags, A, K, T, strict):
    """The second cost functor: the final state charged by a Cost object of its own
    (mppi_set_terminal_weights) through every rollout kernel variant, costs bit exact against the
    oracle, whose terminal-weight mode tests/test_oracle.py recomposes from the reference's own
    Cost class."""
    import mppi_gpu_b200 as m
    cfg = REF_CFG[A]
    wf = np.random.default_rng(A * 31 + K).uniform(2.0, 60.0, 2 * A).astype(np.float32)
    x0, U, eps = make_inputs(K, T, A, seed=A * 55 + K, sigma=0.2)
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        ctl = m.PointMassModel(K, T, 0.1, 2 * A, A, flags=flags | (1 if strict else 0), lam=3.0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    ctl.set_terminal_weights(wf)
    ctl.memcpy_set_data(x0, U, cfg["goal"], cfg["w"])      # sticky across set_problem
    ctl.set_noise(eps)
    na = ctl.get_act()
    inf = ctl.get_inf(want_e=False)
    info = ctl.step_info()
    arith = oracle.ARITH_STRICT if strict else oracle.ARITH_FMA
    p = oracle.make_problem(K, T, A, 0.1, cfg["goal"], cfg["w"], lam=3.0, arith=arith, w_final=wf)
    ref = oracle.step(p, x0, U, eps)
    assert np.array_equal(bits(inf["cost"]), bits(ref["S"])), np.abs(inf["cost"] - ref["S"]).max()
    assert info["argmin"] == ref["argmin"] and bits(inf["beta"]) == bits(ref["beta"])
    assert np.allclose(inf["u"].ravel(), ref["U"].ravel(), rtol=1e-5, atol=1e-6)
    assert np.allclose(na, ref["next_act"], rtol=1e-5, atol=1e-6)
    # None returns to the reference's single Cost object
    ctl.set_terminal_weights(None)
    ctl.memcpy_set_data(x0, U, cfg["goal"], cfg["w"])
    ctl.get_act()
    p0 = oracle.make_problem(K, T, A, 0.1, cfg["goal"], cfg["w"], lam=3.0, arith=arith)
    assert np.array_equal(bits(ctl.get_inf(want_e=False)["cost"]), bits(oracle.step(p0, x0, U, eps)["S"]))
    ctl.close()


@pytest.mark.parametrize("flags", [0, 32, 128, 1024], ids=["unfused", "fused", "step_kernel", "tile_kernel"])
def test_terminal_weights_sampled_chains(oracle, flags):
    """Sampled noise through every chain with terminal weights; the state given as (q, q_dot)."""
    import mppi_gpu_b200 as m
    K, T, A = 6000, 40, 3
    cfg = REF_CFG[A]
    wf = np.array([40, 40, 40, 8, 8, 8], np.float32)
    x0, U, _ = make_inputs(K, T, A, seed=23)
    ctl = m.PointMassModel(K, T, 0.1, 2 * A, A, seed=11, flags=flags)
    ctl.memcpy_set_data(np.zeros(2 * A), U, cfg["goal"], cfg["w"])
    ctl.set_terminal_weights(wf)
    ctl.set_q(x0[:A], x0[A:])
    p = oracle.make_problem(K, T, A, 0.1, cfg["goal"], cfg["w"], arith=oracle.ARITH_FMA, w_final=wf)
    for _ in range(2):
        pre = ctl.get_u()
        na = ctl.get_act()
        inf = ctl.get_inf()
        info = ctl.step_info()
        ref = oracle.step(p, x0, pre, inf["e"])
        assert np.array_equal(bits(inf["cost"]), bits(ref["S"]))
        assert info["argmin"] == ref["argmin"] and bits(inf["beta"]) == bits(ref["beta"])
        assert np.allclose(inf["u"].ravel(), ref["U"].ravel(), rtol=1e-5, atol=1e-6)
        assert np.allclose(na, ref["next_act"], rtol=1e-5, atol=1e-6)
    ctl.close()
