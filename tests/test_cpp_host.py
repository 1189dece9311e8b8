"""C++ host side over the C ABI (cpp/): the reference's driver loop with the shim class.

CPU: the config reader against the reference's own parser self-test constants
(verify_parse, src/main.cu:686-725).  GPU: the closed loop of src/main.cu:326-374 with the
stand-in plant, 200 control steps, trajectory CSV in the reference's format."""
import os
import subprocess

import pytest

from conftest import ROOT

BIN = os.path.join(ROOT, "cpp", "mppi_main")


def _build():
    subprocess.run(["make", "-C", os.path.join(ROOT, "cpp")], check=True, capture_output=True)


def test_config_reader_known_answers():
    _build()
    r = subprocess.run([BIN, "-c", os.path.join(ROOT, "config", "mppi-config-test.yaml"),
                        "--verify-config"], capture_output=True, text=True)
    assert r.returncode == 0 and "Test passed" in r.stdout, r.stdout + r.stderr
    assert "N 3 steps: 12 State dim: 4" in r.stdout


def test_missing_key_exits_like_the_reference(tmp_path):
    _build()
    cfg = tmp_path / "bad.yaml"
    cfg.write_text("---\nenv: e.xml\nsamples: 10\nstate-dim: 2\naction-dim: 1\ndt: 0.1\n")
    r = subprocess.run([BIN, "-c", str(cfg), "--verify-config"], capture_output=True, text=True)
    assert r.returncode == 1
    assert "Please provide the prediction horizon in the config file" in r.stdout


def test_host_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _build()
    r = subprocess.run([BIN, "-c", os.path.join(ROOT, "config", "point_mass2d.yaml"), "--steps", "1"],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "API error failed" in r.stdout     # CUDA_CALL_CONST convention


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,plant", [("point_mass2d", "ideal"), ("point_mass2d", "mjcf"),
                                       ("point_mass1d", "ideal"), ("point_mass3d", "ideal")])
def test_closed_loop_driver(tmp_path, cfg, plant):
    _build()
    traj = tmp_path / "traj.csv"
    r = subprocess.run([BIN, "-c", os.path.join(ROOT, "config", cfg + ".yaml"), "--samples", "20000",
                        "--horizon", "50", "--steps", "200", "--plant", plant, "--quiet",
                        "--honour-config", "-t", str(traj)], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Average controller execution time" in r.stdout
    final = [float(v) for v in r.stdout.split("final state:")[1].split("\n")[0].split()]
    A = len(final) // 2
    goal = {1: [1], 2: [1, 0], 3: [1, .5, .75]}[A]
    if plant == "ideal":
        # 200 control steps (20 s): the mass has covered most of the way to the goal position
        # (velocity is penalised 5-50x more than position in these configs) and moves slowly
        assert all(abs(p - g) < 0.5 * max(abs(g), 0.3) for p, g in zip(final[:A], goal)), final
        assert all(abs(v) < 0.3 for v in final[A:]), final
    else:
        # MJCF-like body: 18.7 m/s^2 per unit control over 0.02 s per control step -- the
        # reference's own plant/model mismatch (src/model_missmatch.cpp); stays in joint range
        assert all(abs(p) <= 1.4 + 1e-6 for p in final[:A]), final
    lines = traj.read_text().strip().split("\n")
    assert lines[0].endswith("size_x,size_u")
    if A == 2:
        assert lines[0] == "x,y,vx,vy,ux,uy,size_x,size_u"       # src/main.cu:41-42
    assert lines[1].split(",")[-2:] == ["201", "200"]
    assert len(lines) == 1 + 200 + 1


@pytest.mark.gpu
def test_driver_with_the_matched_model():
    """--model mjcf: the controller plans with the MJCF body's own dynamics
    (MPPI_MODEL_LINEAR_AXIS) instead of the double integrator; the loop runs, the mass moves
    towards the goal and stays inside the joint range."""
    _build()
    r = subprocess.run([BIN, "-c", os.path.join(ROOT, "config", "point_mass2d.yaml"), "--samples",
                        "20000", "--horizon", "50", "--steps", "300", "--plant", "mjcf", "--model",
                        "mjcf", "--quiet", "--honour-config"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    final = [float(v) for v in r.stdout.split("final state:")[1].split("\n")[0].split()]
    assert 0.0 < final[0] <= 1.4 and abs(final[1]) <= 1.4, final


@pytest.mark.gpu
def test_matched_model_predicts_the_mjcf_body():
    """The gains `mppi_main --model mjcf` passes mean what they say: the controller's predicted
    trajectories (get_inf tap) follow the exact solution of m q'' = gear u - damping q' under the
    same piecewise-constant controls, while the reference's double integrator (dt = 0.1, unit
    gain) is far off -- the reference's model-mismatch experiment (src/model_missmatch.cpp)."""
    import math
    import numpy as np
    import mppi_gpu_b200 as m
    from conftest import make_inputs
    K, T, A, h = 64, 50, 2, 0.02
    mass, gear, damping = 1000.0 * 4.0 / 3.0 * math.pi * 0.05 ** 3 + 0.01, 10.0, 0.1
    k = damping / mass
    g = [1.0, h - 0.5 * h * h * k, 0.0, 1.0 - h * k]
    b = [0.5 * h * h * gear / mass, h * gear / mass]
    x0, U, eps = make_inputs(K, T, A, seed=4, sigma=0.05, u_scale=0.3)
    exact = np.zeros((K, T + 1, 2 * A))
    exact[:, 0] = x0
    ek = math.exp(-k * h)
    for t in range(T):
        u = U[t][None, :].astype(np.float64) + eps[:, t].astype(np.float64)
        p, v = exact[:, t, :A], exact[:, t, A:]
        vt = gear * u / damping                          # terminal velocity under this control
        exact[:, t + 1, A:] = vt + (v - vt) * ek
        exact[:, t + 1, :A] = p + vt * h + (v - vt) * (1 - ek) / k
    err = {}
    for name, kw in (("matched", dict(state_gain=g, act_gain=b)), ("double_integrator", {})):
        ctl = m.PointMassModel(K, T, 0.1, 2 * A, A, **kw)
        ctl.memcpy_set_data(x0, U, [1, 0, 0, 0], [1, 1, 50, 50])
        ctl.set_noise(eps)
        ctl.get_act()
        x = ctl.get_inf(want_x=True, want_e=False)["x"].astype(np.float64)
        ctl.close()
        err[name] = np.abs(x - exact).max()
    assert err["matched"] < 2e-3 * np.abs(exact).max() + 1e-4, err
    assert err["double_integrator"] > 20 * err["matched"], err


@pytest.mark.gpu
def test_step_dump_matches_reference_csv_format_and_reproduces_the_update(tmp_path):
    """-s: the reference's per-step dump (to_csv2, src/main.cu:90-156) that its
    scripts/plot_csv.py consumes; the NumPy recomputation in that script
    (scripts/plot_csv.py:77-108: cost, beta, exp, eta, weights) must reproduce the dumped
    weights from the dumped costs."""
    import csv

    import numpy as np
    _build()
    prefix = tmp_path / "step"
    r = subprocess.run([BIN, "-c", os.path.join(ROOT, "config", "point_mass2d.yaml"), "--samples", "300",
                        "--horizon", "20", "--steps", "2", "--quiet", "-s", str(prefix)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rows = list(csv.DictReader(open(str(prefix) + "1")))
    assert list(rows[0].keys()) == ["sample", "x", "y", "x_dot", "y_dot", "e_x", "e_y", "u[0]", "u[1]",
                                    "u_prev[0]", "u_prev[1]", "c", "w"]
    K, T = 300, 20
    assert len(rows) == K * (T + 1)
    c = np.array([float(rows[n]["c"]) for n in range(K)])
    w = np.array([float(rows[n]["w"]) for n in range(K)])
    ex = np.exp(-(c - c.min()))                       # lambda = 1, plot_csv.py:90-100
    # the CSV carries 6 significant digits (default ostream precision, as the reference's)
    assert np.allclose(w, ex / ex.sum(), rtol=5e-3, atol=1e-9)
    # U_next[t] = U_prev[t+1] + sum_k w_k e_k[t+1]   (update then shift)
    e = np.array([[float(rows[k * (T + 1) + j]["e_x"]) for j in range(T)] for k in range(K)])
    u_prev = np.array([float(rows[j]["u_prev[0]"]) for j in range(T)])
    u_new = np.array([float(rows[j]["u[0]"]) for j in range(T)])
    upd = u_prev + (w[:, None] * e).sum(0)
    assert np.allclose(u_new[:-1], upd[1:], rtol=1e-3, atol=1e-5)
    assert np.isclose(u_new[-1], upd[-1], rtol=1e-3, atol=1e-5)
