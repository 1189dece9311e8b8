"""Generate tests/golden/*.npz by RUNNING THE REFERENCE'S OWN SOURCES.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py

For every case the inputs are written out in full (no dependence on an RNG
implementation at test time) together with
  S_ref     rollout costs from oracle/_ref == the reference's point_mass_gpu.cu +
            cost.cu compiled for the host (PointMassModelGpu::init/run, Cost::*)
  x_ref     trajectories from the same run (small cases only)
and, derived from S_ref with the reference's formulae as restated in
oracle/mppi_oracle.c (src/point_mass.cu:518,751; src/test.cu:97-105;
src/point_mass.cu:805-824):
  beta, argmin, eta, weights, U_next (post-shift), next_act
plus S_fma: the same rollouts with the FMA contraction nvcc applies to the
reference's device build (oracle ORACLE_ARITH_FMA).

The reference has no golden vectors of its own for this path; its
closed-form unit-test fixtures (src/test.cu:11-59,77-105) are regenerated on
the fly in tests/test_oracle.py instead of being stored.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import pyoracle as po  # noqa: E402

# (name, K, T, A, dt, goal, w, x0 scale, U scale, sigma) -- shapes from the reference:
#   cfgtest: config/mppi-config-test.yaml (K=3, T=12, A=2, its goal / cost.w)
#   testcu : src/test.cu MAX_N-1=59, MAX_T-1=99, MAX_A=2
#   pm1d/2d/3d: config/point_mass{1,2,3}d.yaml goal / cost.w / T=50, K cut to 512
CASES = [
    ("cfgtest", 3, 12, 2, 0.1, [1, 2, 3, 4], [1, 2, 0.5, 0.75], 0.0, 0.1, 0.25),
    ("testcu", 59, 99, 2, 0.1, [1, 0, 0, 0], [1, 1, 50, 50], 0.1, 0.2, 0.025),
    ("pm1d", 512, 50, 1, 0.1, [1, 0], [1, 5], 0.05, 0.1, 0.025),
    ("pm2d", 512, 50, 2, 0.1, [1, 0, 0, 0], [1, 1, 50, 50], 0.05, 0.1, 0.025),
    ("pm3d", 512, 50, 3, 0.1, [1, 0.5, 0.75, 0, 0, 0], [1, 1, 1, 5, 5, 5], 0.05, 0.1, 0.025),
    ("pm3d_ragged", 257, 37, 3, 0.05, [1, 0.5, 0.75, 0, 0, 0], [1, 1, 1, 5, 5, 5], 0.2, 0.3, 0.1),
    # caller-given gains (PointMassModelGpu::init takes them as arguments,
    # src/point_mass_gpu.cu:25-39): the damped, geared point mass of envs/point_mass2d.xml
    # (mass 0.5236 + armature 0.01, gear 10, damping 0.1) discretised at dt = 0.1, and a
    # generic non-trivial set with every gain active
    ("pm2d_damped", 384, 60, 2, 0.1, [1, 0, 0, 0], [1, 1, 50, 50], 0.05, 0.1, 0.025),
    ("pm3d_gains", 200, 45, 3, 0.1, [1, 0.5, 0.75, 0, 0, 0], [1, 1, 1, 5, 5, 5], 0.1, 0.2, 0.05),
]


def damped_gains(dt, mass=0.5236 + 0.01, gear=10.0, damping=0.1):
    """p' = p + dt v + dt^2/2 a, v' = v + dt a with a = (gear (u+e) - damping v) / mass."""
    k = damping / mass
    g = np.array([1.0, dt - 0.5 * dt * dt * k, 0.0, 1.0 - dt * k], np.float32)
    b = np.array([0.5 * dt * dt * gear / mass, dt * gear / mass], np.float32)
    return g, b


GAINS = {
    "pm2d_damped": damped_gains(0.1),
    "pm3d_gains": (np.array([0.98, 0.11, -0.02, 0.93], np.float32), np.array([0.007, 0.12], np.float32)),
}


def main():
    assert po.ref_available(), "oracle/_ref is missing: run `make -C oracle` first"
    for i, (name, K, T, A, dt, goal, w, xs, us, sig) in enumerate(CASES):
        rs = np.random.RandomState(1000 + i)
        eps = (sig * rs.standard_normal((K, T, A))).astype(np.float32)
        U = (us * rs.standard_normal((T, A))).astype(np.float32)
        x0 = (xs * rs.standard_normal(2 * A)).astype(np.float32)
        goal = np.asarray(goal, np.float32)
        w = np.asarray(w, np.float32)
        small = K * (T + 1) * 2 * A <= 40000
        gains = GAINS.get(name)
        out = po.ref_rollout_all(K, T, A, dt, 1.0, x0, U, goal, w, eps, want_traj=small, gains=gains)
        S_ref, x_ref = out if small else (out, None)

        p = po.make_problem(K, T, A, dt, goal, w, gains=gains)
        S_orc = po.rollout_all(p, x0, U, eps)
        assert np.array_equal(S_orc.view(np.uint32), S_ref.view(np.uint32)), name
        r = po.step(p, x0, U, eps)
        pf = po.make_problem(K, T, A, dt, goal, w, arith=po.ARITH_FMA, gains=gains)
        S_fma = po.rollout_all(pf, x0, U, eps)
        rf = po.step(pf, x0, U, eps)

        d = dict(K=K, T=T, A=A, dt=np.float32(dt), lam=np.float32(1.0), goal=goal, w=w, x0=x0,
                 U=U, eps=eps, S_ref=S_ref, S_fma=S_fma, beta=r["beta"], argmin=r["argmin"],
                 eta=r["eta"], weights=r["weights"], U_next=r["U"], next_act=r["next_act"],
                 U_next_fma=rf["U"], next_act_fma=rf["next_act"], beta_fma=rf["beta"],
                 argmin_fma=rf["argmin"], eta_fma=rf["eta"])
        if small:
            d["x_ref"] = x_ref
        if gains is not None:
            d["state_gain"], d["act_gain"] = gains
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **d)
        print(f"{name}: K={K} T={T} A={A} beta={r['beta']:.6f} eta={r['eta']:.6f} "
              f"argmin={r['argmin']} -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
