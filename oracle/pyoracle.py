"""ctypes bindings for the CPU checker (oracle/liboracle.so, oracle/_ref/libmppi_ref.so).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg; never from mppi_gpu_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MAX_ACT = 8
ARITH_STRICT, ARITH_FMA = 0, 1

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


class OracleProblem(C.Structure):
    _fields_ = [
        ("K", C.c_int32), ("T", C.c_int32), ("A", C.c_int32), ("arith", C.c_int32),
        ("dt", C.c_float), ("lambda_", C.c_float),
        ("inv_s", C.c_float * MAX_ACT),
        ("goal", C.c_float * (2 * MAX_ACT)),
        ("w", C.c_float * (2 * MAX_ACT)),
        ("use_gains", C.c_int32),
        ("g", C.c_float * 4),
        ("b", C.c_float * 2),
        ("use_wf", C.c_int32),
        ("wf", C.c_float * (2 * MAX_ACT)),
    ]


def make_problem(K, T, A, dt, goal, w, lam=1.0, inv_s=None, arith=ARITH_STRICT, gains=None,
                 w_final=None):
    """gains = (state_gain[4], act_gain[2]) selects caller-given gains (the arguments of the
    reference's PointMassModelGpu::init) instead of the double integrator formed from dt.
    w_final = weights of a second Cost object that charges the final state (Cost::final_cost);
    None: the reference's single object, w."""
    p = OracleProblem()
    if w_final is not None:
        p.use_wf = 1
        for i in range(2 * A):
            p.wf[i] = float(w_final[i])
    if gains is not None:
        p.use_gains = 1
        for i in range(4):
            p.g[i] = float(gains[0][i])
        for i in range(2):
            p.b[i] = float(gains[1][i])
    p.K, p.T, p.A, p.arith = int(K), int(T), int(A), int(arith)
    p.dt, p.lambda_ = float(dt), float(lam)
    inv_s = [1.0] * A if inv_s is None else list(inv_s)
    for i in range(A):
        p.inv_s[i] = float(inv_s[i])
    for i in range(2 * A):
        p.goal[i] = float(goal[i])
        p.w[i] = float(w[i])
    return p


def build(ref: bool = True) -> None:
    """(Re)build liboracle.so and, when /root/reference is present, oracle/_ref."""
    subprocess.run(["make", "-C", HERE, "all" if ref else os.path.join(HERE, "liboracle.so")],
                   check=True, capture_output=True)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = C.CDLL(path)
        PP = C.POINTER(OracleProblem)
        L.oracle_gains.argtypes = [C.c_float, _f32p, _f32p]
        L.oracle_rollout.argtypes = [PP, _f32p, _f32p, _f32p, C.c_void_p]
        L.oracle_rollout.restype = C.c_float
        L.oracle_rollout_all.argtypes = [PP, _f32p, _f32p, _f32p, _f32p, C.c_void_p, C.c_int]
        L.oracle_beta.argtypes = [_f32p, C.c_int64, C.POINTER(C.c_int64)]
        L.oracle_beta.restype = C.c_float
        L.oracle_exp.argtypes = [_f32p, C.c_int64, C.c_float, C.c_float, _f32p]
        L.oracle_eta.argtypes = [_f32p, C.c_int64, C.POINTER(C.c_double)]
        L.oracle_eta.restype = C.c_float
        L.oracle_weights.argtypes = [_f32p, C.c_int64, C.c_float, C.c_float, C.c_float, _f32p]
        L.oracle_update_act.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int]
        L.oracle_update_act_f64.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int]
        L.oracle_update_act_mt.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.oracle_shift.argtypes = [_f32p, C.c_int, C.c_int]
        L.oracle_step.argtypes = [PP, _f32p, _f32p, _f32p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                           C.POINTER(C.c_uint32)]
        L.oracle_sample_eps.argtypes = [C.c_uint64, C.c_uint64, C.c_int64, C.c_int64, C.c_int,
                                        C.c_int, _f32p, _f32p]
        L.oracle_philox4x32.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int,
                                        C.POINTER(C.c_uint32)]
        L.oracle_sample_eps_rounds.argtypes = [C.c_uint64, C.c_uint64, C.c_int64, C.c_int64, C.c_int,
                                               C.c_int, _f32p, C.c_int, _f32p]
        _lib = L
    return _lib


def ref_available() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libmppi_ref.so"))


def ref():
    """The reference's own model/cost sources compiled for the host (oracle/_ref)."""
    global _ref
    if _ref is None:
        path = os.path.join(HERE, "_ref", "libmppi_ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle` where /root/reference exists)")
        L = C.CDLL(path)
        L.ref_rollout_costs.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _f32p,
                                        _f32p, _f32p, _f32p, _f32p, _f32p, C.c_void_p, C.c_int]
        L.ref_rollout_costs.restype = C.c_int
        L.ref_rollout_costs_gains.argtypes = [C.c_int, C.c_int, C.c_int, _f32p, _f32p, C.c_float,
                                              _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_void_p,
                                              C.c_int]
        L.ref_rollout_costs_gains.restype = C.c_int
        L.ref_cost_terms.argtypes = [C.c_int, C.c_int, _f32p, _f32p, C.c_float, _f32p, _f32p, _f32p,
                                     _f32p, _f32p, _f32p]
        L.ref_cost_terms.restype = C.c_int
        _ref = L
    return _ref


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# --------------------------------------------------------------------------- wrappers
def gains(dt):
    g = np.zeros(4, np.float32)
    b = np.zeros(2, np.float32)
    lib().oracle_gains(float(dt), g, b)
    return g, b


def rollout_all(p, x0, U, eps, want_traj=False, nthreads=1):
    S = np.zeros(p.K, np.float32)
    xt = np.zeros((p.K, p.T + 1, 2 * p.A), np.float32) if want_traj else None
    lib().oracle_rollout_all(C.byref(p), _f32(x0), _f32(U), _f32(eps), S,
                             xt.ctypes.data if xt is not None else None, int(nthreads))
    return (S, xt) if want_traj else S


def ref_rollout_all(K, T, A, dt, lam, x0, U, goal, w, eps, want_traj=False, nthreads=1,
                    gains=None):
    S = np.zeros(K, np.float32)
    xt = np.zeros((K, T + 1, 2 * A), np.float32) if want_traj else None
    e = _f32(eps).copy()
    if gains is not None:
        rc = ref().ref_rollout_costs_gains(K, T, A, _f32(gains[0]), _f32(gains[1]), float(lam),
                                           _f32(x0), _f32(U), _f32(goal), _f32(w), e, S,
                                           xt.ctypes.data if xt is not None else None,
                                           int(nthreads))
        assert rc == 0
        return (S, xt) if want_traj else S
    rc = ref().ref_rollout_costs(K, T, A, float(dt), float(lam), _f32(x0), _f32(U), _f32(goal),
                                 _f32(w), e, S, xt.ctypes.data if xt is not None else None,
                                 int(nthreads))
    assert rc == 0
    return (S, xt) if want_traj else S


def ref_cost_terms(A, w, goal, lam, inv_s, x, u, e):
    """The reference's own Cost class (oracle/_ref): (step_cost, final_cost) of n triples
    x [n,2A], u [n,A], e [n,A] for the weights w."""
    x, u, e = _f32(x).reshape(-1, 2 * A), _f32(u).reshape(-1, A), _f32(e).reshape(-1, A)
    n = x.shape[0]
    st, fi = np.zeros(n, np.float32), np.zeros(n, np.float32)
    rc = ref().ref_cost_terms(n, int(A), _f32(w), _f32(goal), float(lam), _f32(inv_s), x, u, e, st, fi)
    assert rc == 0
    return st, fi


def beta(S):
    idx = C.c_int64(-1)
    b = lib().oracle_beta(_f32(S), len(S), C.byref(idx))
    return np.float32(b), int(idx.value)


def exp(S, lam, b):
    out = np.zeros(len(S), np.float32)
    lib().oracle_exp(_f32(S), len(S), float(lam), float(b), out)
    return out


def eta(ex):
    d = C.c_double(0)
    s = lib().oracle_eta(_f32(ex), len(ex), C.byref(d))
    return np.float32(s), float(d.value)


def weights(S, lam, b, e):
    out = np.zeros(len(S), np.float32)
    lib().oracle_weights(_f32(S), len(S), float(lam), float(b), float(e), out)
    return out


def update_act(u, w, e, n, t, a, f64=False, nthreads=1):
    u = _f32(u).copy()
    if nthreads > 1 and not f64:
        lib().oracle_update_act_mt(u, _f32(w), _f32(e), int(n), int(t), int(a), int(nthreads))
        return u
    fn = lib().oracle_update_act_f64 if f64 else lib().oracle_update_act
    fn(u, _f32(w), _f32(e), int(n), int(t), int(a))
    return u


def shift(u, T, A):
    u = _f32(u).copy()
    lib().oracle_shift(u, int(T), int(A))
    return u


def step(p, x0, U, eps, nthreads=1):
    """One reference control step.  Returns dict(U, next_act, S, beta, eta, weights, argmin)."""
    U = _f32(U).copy()
    next_act = np.zeros(p.A, np.float32)
    S = np.zeros(p.K, np.float32)
    w = np.zeros(p.K, np.float32)
    b = C.c_float(0)
    e = C.c_float(0)
    am = C.c_int64(-1)
    lib().oracle_step(C.byref(p), _f32(x0), U, _f32(eps), next_act.ctypes.data, S.ctypes.data,
                      C.addressof(b), C.addressof(e), w.ctypes.data, C.addressof(am),
                      int(nthreads))
    return dict(U=U, next_act=next_act, S=S, beta=np.float32(b.value), eta=np.float32(e.value),
                weights=w, argmin=int(am.value))


def philox(ctr, key, rounds=10):
    c = (C.c_uint32 * 4)(*[int(x) & 0xFFFFFFFF for x in ctr])
    k = (C.c_uint32 * 2)(*[int(x) & 0xFFFFFFFF for x in key])
    o = (C.c_uint32 * 4)()
    lib().oracle_philox4x32(c, k, int(rounds), o)
    return [int(x) for x in o]


def sample_eps(seed, step_idx, k0, K, T, A, sigma, rounds=10):
    eps = np.zeros((K, T, A), np.float32)
    lib().oracle_sample_eps_rounds(int(seed), int(step_idx), int(k0), int(K), int(T), int(A),
                                   _f32(sigma), int(rounds), eps)
    return eps
