"""Pipelined sampling (MPPI_FLAG_PIPELINED_SAMPLING, csrc/controller.cu mppi_step_enqueue): the
noise of step n+1 is drawn on a second stream, into a second eps buffer, while step n runs.

The sampler is counter based -- eps depends on (seed, step, k, t, a) only -- and the chain that
consumes the noise is the same kernels in the same order, so every result must equal the plain
three-kernel chain's BIT FOR BIT: noise, rollout costs, beta, argmin, eta, U, next action.  The
plain chain itself is held to the oracle by tests/test_gpu_parity.py; one case here repeats that
check directly."""
import numpy as np
import pytest

from conftest import REF_CFG, bits, make_inputs
from test_gpu_parity import _assert_parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    import mppi_gpu_b200 as m
    return m


def _pair(M, K, T, A, seed, extra=0):
    from mppi_gpu_b200 import capi
    cfg = REF_CFG[A]
    x0, U, _ = make_inputs(K, T, A, seed=seed)
    ctls = []
    for fl in (0, capi.FLAG_PIPELINED_SAMPLING):
        c = M.PointMassModel(K, T, 0.1, 2 * A, A, seed=seed + 1, flags=fl | extra)
        c.memcpy_set_data(x0, U, cfg["goal"], cfg["w"])
        ctls.append(c)
    assert ctls[1].flags() & capi.FLAG_PIPELINED_SAMPLING
    return ctls, x0


def _same_state(plain, pipe, want_e=True):
    a, b = plain.get_inf(want_e=want_e), pipe.get_inf(want_e=want_e)
    for k in ("cost", "u") + (("e",) if want_e else ()):
        assert np.array_equal(bits(a[k]), bits(b[k])), k
    assert bits(a["beta"]) == bits(b["beta"]) and bits(a["nabla"]) == bits(b["nabla"])
    assert plain.step_info()["argmin"] == pipe.step_info()["argmin"]


# ragged K (pad columns), every A, T*A not a multiple of the 40-row TMA box, TMA and
# register-pipelined rollout variants (small and large K)
@pytest.mark.parametrize("A,K,T", [(2, 3000, 50), (1, 1027, 37), (3, 40001, 20), (4, 777, 41),
                                   (3, 700000, 4)])
def test_pipelined_chain_equals_plain_chain_bitwise(M, A, K, T):
    (plain, pipe), x0 = _pair(M, K, T, A, seed=11)
    rs = np.random.RandomState(3)
    for step in range(5):
        na0, na1 = plain.get_act(), pipe.get_act()
        assert np.array_equal(bits(na0), bits(na1)), f"next action, step {step}"
        _same_state(plain, pipe, want_e=(K * T * A < 3e6 or step == 4))
        x = (x0 + 0.01 * rs.standard_normal(2 * A)).astype(np.float32)
        plain.set_x(x)
        pipe.set_x(x)
    # one sampler in line before the first step, then {rollout, average, sampler ahead} per step
    assert pipe.launch_count() == 1 + 5 * 3
    assert plain.launch_count() == 5 * 3
    plain.close()
    pipe.close()


def test_pipelined_step_matches_oracle(M, oracle):
    """The third pipelined step (both buffers have been through a swap) against the oracle on
    the noise the GPU drew."""
    from mppi_gpu_b200 import capi
    K, T, A = 3000, 50, 2
    cfg = REF_CFG[A]
    x0, U, _ = make_inputs(K, T, A, seed=21)
    ctl = M.PointMassModel(K, T, 0.1, 2 * A, A, seed=9, flags=capi.FLAG_PIPELINED_SAMPLING)
    ctl.memcpy_set_data(x0, U, cfg["goal"], cfg["w"])
    for _ in range(2):
        ctl.get_act()
    U_before = ctl.get_u()
    na = ctl.get_act()
    inf = ctl.get_inf()
    p = oracle.make_problem(K, T, A, 0.1, cfg["goal"], cfg["w"], arith=oracle.ARITH_FMA)
    ref = oracle.step(p, x0, U_before.reshape(T, A), inf["e"])
    _assert_parity(na, inf, ctl.step_info(), ref, K, T, A)
    ctl.close()


def test_pipelined_survives_mode_switches(M):
    """Injected-noise steps, profiling steps, sample_only and set_problem in between: noise drawn
    ahead is dropped whenever a step of another kind advanced the counter, and drawn again."""
    K, T, A = 2049, 23, 2
    (plain, pipe), x0 = _pair(M, K, T, A, seed=5)
    _, _, eps = make_inputs(K, T, A, seed=77)
    cfg = REF_CFG[A]

    def both(f):
        return f(plain), f(pipe)

    def check(tag):
        na0, na1 = both(lambda c: c.get_act())
        assert np.array_equal(bits(na0), bits(na1)), tag
        _same_state(plain, pipe)

    # leave the pipeline after an odd and after an even number of pipelined steps: the plain
    # graphs are bound to one eps buffer, the pipeline alternates between two
    check("sampled 0")
    both(lambda c: c.set_noise(eps))                # switches to injected mode
    check("injected after 1 pipelined step")
    both(lambda c: c.set_noise_mode(False))
    check("sampled after injected")                 # must re-prime: the counter moved on
    check("sampled, pipelined again")
    both(lambda c: c.set_noise(eps[::-1].copy()))
    check("injected after 2 pipelined steps")
    check("injected again")
    both(lambda c: c.set_noise_mode(False))
    check("sampled 1")
    check("sampled 2")
    check("sampled 3")
    both(lambda c: c.set_noise(eps))
    check("injected after 3 pipelined steps")
    both(lambda c: c.set_noise_mode(False))
    check("sampled 4")
    both(lambda c: c.set_profiling(True))
    check("profiled step (plain chain)")
    both(lambda c: c.set_profiling(False))
    check("sampled after profiling")
    both(lambda c: c.sample_only(1234))             # overwrites the current buffer
    a, b = both(lambda c: c.get_inf())
    assert np.array_equal(bits(a["e"]), bits(b["e"]))
    check("sampled after sample_only")
    x0b, Ub, _ = make_inputs(K, T, A, seed=8)
    both(lambda c: c.memcpy_set_data(x0b, Ub, cfg["goal"], cfg["w"]))
    check("after set_problem")
    check("and again")
    plain.close()
    pipe.close()


def test_pipelined_flag_is_ignored_where_it_does_not_apply(M):
    """With fused sampling or the one-kernel step there is no separate sampler to move."""
    from mppi_gpu_b200 import capi
    K, T, A = 3000, 50, 2
    cfg = REF_CFG[A]
    x0, U, _ = make_inputs(K, T, A, seed=2)
    outs = []
    for fl in (capi.FLAG_FUSED_SAMPLING, capi.FLAG_FUSED_SAMPLING | capi.FLAG_PIPELINED_SAMPLING):
        c = M.PointMassModel(K, T, 0.1, 2 * A, A, seed=4, flags=fl)
        c.memcpy_set_data(x0, U, cfg["goal"], cfg["w"])
        na = [c.get_act() for _ in range(3)][-1]
        outs.append((na, c.get_u(), c.launch_count()))
        c.close()
    assert np.array_equal(bits(outs[0][0]), bits(outs[1][0]))
    assert np.array_equal(bits(outs[0][1]), bits(outs[1][1]))
    assert outs[0][2] == outs[1][2] == 3 * 2
