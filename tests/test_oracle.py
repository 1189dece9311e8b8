"""CPU tests: pin the oracle (oracle/mppi_oracle.c) to the reference.

 * bit-for-bit against the reference's own model/cost sources compiled for the host
   (oracle/_ref), on the shapes of the reference's configs;
 * against the committed golden vectors (tests/golden, produced by that build);
 * against the closed-form fixtures of the reference's unit tests
   (src/test.cu:11-59 test_exp, :77-105/:181-229 test_update_act);
 * Philox-4x32-10 against the Random123 known-answer vectors.
"""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, REF_CFG, bits, golden_gains, load_golden, make_inputs


# ------------------------------------------------------------------ vs oracle/_ref
@pytest.mark.parametrize("A", [1, 2, 3])
@pytest.mark.parametrize("K,T", [(3000, 50), (3, 12), (59, 99), (1, 1), (10000, 200)])
def test_oracle_matches_ref_bit_exact(oracle, A, K, T):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    if K == 10000 and A != 3:
        pytest.skip("one README-shape case is enough")
    cfg = REF_CFG[A]
    x0, U, eps = make_inputs(K, T, A, seed=K + 7 * T + A)
    want_traj = K * T <= 160000
    ref = oracle.ref_rollout_all(K, T, A, 0.1, 1.0, x0, U, cfg["goal"], cfg["w"], eps,
                                 want_traj=want_traj, nthreads=4)
    p = oracle.make_problem(K, T, A, 0.1, cfg["goal"], cfg["w"])
    got = oracle.rollout_all(p, x0, U, eps, want_traj=want_traj, nthreads=4)
    if want_traj:
        assert np.array_equal(bits(got[0]), bits(ref[0]))
        assert np.array_equal(got[1], ref[1])       # value-equal (sign of zero aside)
    else:
        assert np.array_equal(bits(got), bits(ref))


def test_oracle_lambda_reaches_ref_cost(oracle):
    """lambda multiplies the control-cost term (src/cost.cu:48); check a non-unit value."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built")
    K, T, A = 64, 20, 2
    cfg = REF_CFG[A]
    x0, U, eps = make_inputs(K, T, A, seed=5, u_scale=1.0)
    ref = oracle.ref_rollout_all(K, T, A, 0.1, 1.5, x0, U, cfg["goal"], cfg["w"], eps)
    p = oracle.make_problem(K, T, A, 0.1, cfg["goal"], cfg["w"], lam=1.5)
    assert np.array_equal(bits(oracle.rollout_all(p, x0, U, eps)), bits(ref))


# ------------------------------------------------------------------ vs golden vectors
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_golden(oracle, name):
    g = load_golden(name)
    K, T, A = int(g["K"]), int(g["T"]), int(g["A"])
    p = oracle.make_problem(K, T, A, float(g["dt"]), g["goal"], g["w"], lam=float(g["lam"]),
                            gains=golden_gains(g))
    S, xt = oracle.rollout_all(p, g["x0"], g["U"], g["eps"], want_traj=True)
    assert np.array_equal(bits(S), bits(g["S_ref"]))
    if "x_ref" in g:
        assert np.array_equal(xt, g["x_ref"])
    r = oracle.step(p, g["x0"], g["U"], g["eps"])
    assert r["argmin"] == int(g["argmin"])
    assert bits(r["beta"]) == bits(g["beta"])
    assert bits(r["eta"]) == bits(g["eta"])
    assert np.array_equal(bits(r["weights"]), bits(g["weights"]))
    assert np.array_equal(bits(r["U"]), bits(g["U_next"]))
    assert np.array_equal(bits(r["next_act"]), bits(g["next_act"]))
    pf = oracle.make_problem(K, T, A, float(g["dt"]), g["goal"], g["w"], lam=float(g["lam"]),
                             arith=oracle.ARITH_FMA, gains=golden_gains(g))
    assert np.array_equal(bits(oracle.rollout_all(pf, g["x0"], g["U"], g["eps"])), bits(g["S_fma"]))


@pytest.mark.parametrize("A,K,T", [(1, 300, 40), (2, 512, 50), (3, 257, 37)])
def test_oracle_general_gains_match_ref_bit_exact(oracle, A, K, T):
    """Caller-given gains (the LinearAxis dynamics functor): PointMassModelGpu::init takes the
    gains as arguments (src/point_mass_gpu.cu:25-39), so the reference's own step() pins them."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    cfg = REF_CFG[A]
    x0, U, eps = make_inputs(K, T, A, seed=3 * K + A, sigma=0.1)
    gains = (np.array([0.97, 0.12, 0.03, 0.9], np.float32), np.array([0.02, 0.3], np.float32))
    ref = oracle.ref_rollout_all(K, T, A, 0.1, 1.0, x0, U, cfg["goal"], cfg["w"], eps,
                                 want_traj=True, nthreads=4, gains=gains)
    p = oracle.make_problem(K, T, A, 0.1, cfg["goal"], cfg["w"], gains=gains)
    got = oracle.rollout_all(p, x0, U, eps, want_traj=True, nthreads=4)
    assert np.array_equal(bits(got[0]), bits(ref[0]))
    assert np.array_equal(got[1], ref[1])


@pytest.mark.parametrize("A,K,T", [(1, 100, 40), (2, 128, 50), (3, 64, 37)])
def test_oracle_terminal_weights_match_the_reference_cost_class(oracle, A, K, T):
    """Terminal weights: a second Cost object charges the final state.  The reference uses ONE
    object for stage and final cost (src/point_mass_gpu.cu:107,116), but its Cost class takes the
    weights as an argument (include/cost.hpp:8-14), so the class itself pins the mode: a rollout is
    recomposed, bit for bit, from Cost(w).step_cost over the reference's own trajectory and
    Cost(w_final).final_cost on its last state, accumulated as run() does (`_c += ...`)."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    cfg = REF_CFG[A]
    rng = np.random.default_rng(7 * A + K)
    wf = rng.uniform(2.0, 60.0, 2 * A).astype(np.float32)
    lam = 0.7
    x0, U, eps = make_inputs(K, T, A, seed=5 * K + A, sigma=0.1)
    _, xt = oracle.ref_rollout_all(K, T, A, 0.1, lam, x0, U, cfg["goal"], cfg["w"], eps, want_traj=True)
    p = oracle.make_problem(K, T, A, 0.1, cfg["goal"], cfg["w"], lam=lam, w_final=wf)
    S = oracle.rollout_all(p, x0, U, eps)
    inv_s = np.ones(A, np.float32)
    want = np.zeros(K, np.float32)
    for k in range(K):
        st, _ = oracle.ref_cost_terms(A, cfg["w"], cfg["goal"], lam, inv_s, xt[k, 1:], U, eps[k])
        _, fi = oracle.ref_cost_terms(A, wf, cfg["goal"], lam, inv_s, xt[k, T:], U[:1], eps[k, :1])
        c = np.float32(0)
        for t in range(T):
            c = np.float32(c + st[t])
        want[k] = np.float32(c + fi[0])
    assert np.array_equal(bits(S), bits(want))
    # and with w_final == w the mode is the reference's run()
    p2 = oracle.make_problem(K, T, A, 0.1, cfg["goal"], cfg["w"], lam=lam, w_final=cfg["w"])
    p1 = oracle.make_problem(K, T, A, 0.1, cfg["goal"], cfg["w"], lam=lam)
    assert np.array_equal(bits(oracle.rollout_all(p2, x0, U, eps)), bits(oracle.rollout_all(p1, x0, U, eps)))


def test_fma_and_strict_agree_to_rounding(oracle):
    g = load_golden("pm3d")
    rel = np.abs(g["S_fma"] - g["S_ref"]) / np.abs(g["S_ref"])
    assert rel.max() < 2e-6
    assert int(np.argmin(g["S_fma"])) == int(np.argmin(g["S_ref"]))


# ------------------------------------------------------------------ reference unit-test fixtures
def test_exp_fixture_of_reference_test_cu(oracle):
    """src/test.cu:11-59: cost_i = i, lambda = 1, beta = 0.25, |exp - out| < TOL = 1e-6."""
    for n in range(1, 60):
        cost = np.arange(n, dtype=np.float32)
        out = oracle.exp(cost, 1.0, 0.25)
        want = np.exp(-1.0 * (cost.astype(np.float64) - 0.25))
        assert np.all(np.abs(want - out) < 1e-6)


def _update_act_fixture(n, t, a):
    """init_update_act_data, src/test.cu:77-95"""
    idx = np.arange(n * t * a, dtype=np.float64)
    e = (0.25 * idx).astype(np.float32)
    w = (0.5 * np.arange(n, dtype=np.float64)).astype(np.float32)
    u = (0.75 * np.arange(t * a, dtype=np.float64)).astype(np.float32)
    return u, w, e


@pytest.mark.parametrize("n", [1, 2, 7, 31, 59])
def test_update_act_fixture_of_reference_test_cu(oracle, n):
    """src/test.cu:97-105 update_act_cpu on the closed-form fixture, t in 1..99, a = 2."""
    a = 2
    for t in (1, 2, 13, 50, 99):
        u, w, e = _update_act_fixture(n, t, a)
        got = oracle.update_act(u, w, e, n, t, a)
        # float32 restatement of the serial loop, vectorised over (t,a)
        want = u.copy()
        e2 = e.reshape(n, t * a)
        for k in range(n):
            want = (want + (w[k] * e2[k]).astype(np.float32)).astype(np.float32)
        assert np.array_equal(bits(got), bits(want))
        assert np.array_equal(bits(oracle.update_act(u, w, e, n, t, a, nthreads=3)), bits(want))
        # and the double-accumulated anchor agrees to float rounding of the result
        d = oracle.update_act(u, w, e, n, t, a, f64=True)
        assert np.allclose(got, d, rtol=1e-5)


def test_shift_repeats_last_row(oracle):
    """shift_act, src/point_mass.cu:805-824"""
    T, A = 7, 3
    u = np.arange(T * A, dtype=np.float32)
    s = oracle.shift(u, T, A).reshape(T, A)
    ref = u.reshape(T, A)
    assert np.array_equal(s[:-1], ref[1:])
    assert np.array_equal(s[-1], ref[-1])


def test_weights_formula_and_normalisation(oracle):
    """src/point_mass.cu:518 (float exp_red) and :751 (double-literal weights_kernel)"""
    rs = np.random.RandomState(3)
    S = (100 + 5 * rs.standard_normal(4096)).astype(np.float32)
    b, am = oracle.beta(S)
    assert am == int(np.argmin(S)) and b == S.min()
    for lam in (1.0, 1.5, 10.0):
        ex = oracle.exp(S, lam, b)
        nil = np.float32(-(np.float32(1) / np.float32(lam)))
        want = np.exp((nil * (S - b)).astype(np.float32).astype(np.float64))
        assert np.allclose(ex, want, rtol=3e-7, atol=0)
        eta32, eta64 = oracle.eta(ex)
        assert abs(eta32 - eta64) / eta64 < 1e-5
        w = oracle.weights(S, lam, b, eta32)
        assert abs(float(w.astype(np.float64).sum()) - 1.0) < 1e-5
        assert np.allclose(w, ex / eta32, rtol=1e-6)


def test_beta_ties_pick_lowest_index(oracle):
    S = np.array([3, 1, 2, 1, 1], np.float32)
    assert oracle.beta(S) == (np.float32(1), 1)


def test_step_equals_composition(oracle):
    """oracle_step == sim, beta, exp, nabla, weights, update_act, next_act, shift
    (get_act, src/point_mass.cu:129-203)"""
    K, T, A = 300, 25, 2
    cfg = REF_CFG[A]
    x0, U, eps = make_inputs(K, T, A, seed=11)
    p = oracle.make_problem(K, T, A, 0.1, cfg["goal"], cfg["w"], lam=2.0)
    r = oracle.step(p, x0, U, eps)
    S = oracle.rollout_all(p, x0, U, eps)
    b, am = oracle.beta(S)
    ex = oracle.exp(S, 2.0, b)
    eta32, _ = oracle.eta(ex)
    w = oracle.weights(S, 2.0, b, eta32)
    un = oracle.update_act(U, w, eps, K, T, A)
    assert np.array_equal(bits(r["S"]), bits(S)) and r["argmin"] == am
    assert np.array_equal(bits(r["next_act"]), bits(un.reshape(T, A)[0]))
    assert np.array_equal(bits(r["U"]), bits(oracle.shift(un, T, A)))


# ------------------------------------------------------------------ Philox
def test_philox_known_answer_vectors(oracle):
    """Random123 kat_vectors, philox4x32 with 10 rounds"""
    kats = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kats:
        assert oracle.philox(ctr, key) == want
    # philox4x32 with 7 rounds (mppi_params.philox_rounds = 7), same kat_vectors file
    kats7 = [
        ([0, 0, 0, 0], [0, 0], [0x5f6fb709, 0x0d893f64, 0x4f121f81, 0x4f730a48]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x5207ddc2, 0x45165e59, 0x4d8ee751, 0x8c52f662]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0x4dfccaba, 0x190a87f0, 0xc47362ba, 0xb6b5242a]),
    ]
    for ctr, key, want in kats7:
        assert oracle.philox(ctr, key, rounds=7) == want


def test_sampler_statistics_and_shard_independence(oracle):
    K, T, A = 4096, 25, 2
    sig = [0.025, 0.1]
    e = oracle.sample_eps(7, 3, 0, K, T, A, sig)
    for a in range(A):
        col = e[:, :, a].astype(np.float64).ravel()
        assert abs(col.mean()) < 4 * sig[a] / np.sqrt(col.size)
        assert abs(col.std() / sig[a] - 1) < 0.02
        z = np.sort(col / sig[a])
        from math import erf
        cdf = 0.5 * (1 + np.vectorize(erf)(z / np.sqrt(2)))
        ks = np.max(np.abs(cdf - (np.arange(z.size) + 0.5) / z.size))
        assert ks < 1.63 / np.sqrt(z.size)          # 1 % KS critical value
    # value of eps[k] depends on the global k only
    part = oracle.sample_eps(7, 3, 1000, 96, T, A, sig)
    assert np.array_equal(part, e[1000:1096])
    # different step / seed -> different stream
    assert not np.array_equal(oracle.sample_eps(7, 4, 0, 64, T, A, sig), e[:64])
    assert not np.array_equal(oracle.sample_eps(8, 3, 0, 64, T, A, sig), e[:64])
