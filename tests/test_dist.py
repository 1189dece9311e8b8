"""Multi-rank tests: world_size-2 gloo on CPU (exchange logic), NCCL on >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

WORKER = os.path.join(ROOT, "tests", "dist_worker.py")


def _launch(mode, nproc, port, *extra):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER, mode, *extra]
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=240)


def test_k_shard_exchange_logic_gloo_world2():
    r = _launch("gloo", 2, 29541)
    assert r.returncode == 0 and "GLOO_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_k_shard_exchange_logic_gloo_world3():
    r = _launch("gloo", 3, 29542)
    assert r.returncode == 0 and "GLOO_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_k_shard_online_softmax_merge_gloo_world3():
    """The single-exchange merge of the peer-mailbox path: every shard averages relative to its
    own minimum, the shards' accumulators are rescaled by exp(-(beta_r-beta)/lambda) and summed."""
    r = _launch("gloo_merge", 3, 29547)
    assert r.returncode == 0 and "GLOO_MERGE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.gpu
def test_k_sharded_one_kernel_step():
    """K-shards running the one-kernel step, merged by the single peer-mailbox exchange."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    r = _launch("p2p", min(n, 4), 29548, "128")
    assert r.returncode == 0 and "P2P_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.gpu
def test_k_sharded_tile_kernel():
    """K-shards running the on-chip tile kernel (MPPI_FLAG_TILE_KERNEL = 1024): the last CTA of
    every shard's kernel runs the NVLink exchange itself -- one kernel per step on any number
    of GPUs; three steps against the single-shard oracle, U bit-identical on every rank."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    r = _launch("p2p", min(n, 4), 29549, "1024")
    assert r.returncode == 0 and "P2P_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.gpu
@pytest.mark.parametrize("comm", ["nccl", "p2p"])
def test_k_sharded_controller(comm):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    r = _launch(comm, min(n, 4), 29543 + (comm == "p2p"))
    assert r.returncode == 0 and comm.upper() + "_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.gpu
@pytest.mark.parametrize("comm", ["nccl", "p2p"])
def test_k_sharded_pipelined_sampling(comm):
    """K-shards with MPPI_FLAG_PIPELINED_SAMPLING (512): every rank draws its shard of the next
    step's noise behind the chain; three steps against the single-shard oracle."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    r = _launch(comm, min(n, 4), 29551 + (comm == "p2p"), "512")
    assert r.returncode == 0 and comm.upper() + "_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.gpu
def test_single_process_device_group(oracle):
    """mppi_create_multi: ONE process (the reference's process model, src/main.cu) drives
    several GPUs; K is sharded, the shards exchange through peer memory; the caller sees one
    controller with full-K arrays."""
    import numpy as np
    import torch
    from conftest import REF_CFG, bits, make_inputs
    import mppi_gpu_b200 as m
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    devs = list(range(min(n, 4)))
    K, T, A, lam = 30001, 40, 3, 4.0
    cfg = REF_CFG[A]
    x0, U, eps = make_inputs(K, T, A, seed=12, sigma=0.2)
    ctl = m.PointMassModel(K, T, 0.1, 2 * A, A, lam=lam, seed=5, devices=devs)
    assert ctl.k_local == K and ctl.k_offset == 0
    ctl.memcpy_set_data(x0, U, cfg["goal"], cfg["w"])
    p = oracle.make_problem(K, T, A, 0.1, cfg["goal"], cfg["w"], lam=lam, arith=oracle.ARITH_FMA)
    # injected noise, full-K arrays in and out
    ctl.set_noise(eps)
    na = ctl.get_act()
    inf = ctl.get_inf(want_x=False)
    info = ctl.step_info()
    ref = oracle.step(p, x0, U, eps, nthreads=4)
    assert np.array_equal(bits(inf["cost"]), bits(ref["S"]))
    assert np.array_equal(inf["e"], eps)
    assert info["argmin"] == ref["argmin"] and bits(inf["beta"]) == bits(ref["beta"])
    assert np.allclose(inf["u"].ravel(), ref["U"].ravel(), rtol=1e-5, atol=1e-6)
    assert np.allclose(na, ref["next_act"], rtol=1e-5, atol=1e-6)
    assert abs(float(inf["weight"].astype(np.float64).sum()) - 1) < 1e-4
    # sampled noise: the same stream as a single-GPU controller with the same seed
    ctl.set_noise_mode(False)
    single = m.PointMassModel(K, T, 0.1, 2 * A, A, lam=lam, seed=5, device=0)
    single.memcpy_set_data(x0, ctl.get_u(), cfg["goal"], cfg["w"])
    for _ in range(2):                       # bring the single controller's step counter level
        single.sample_only(0)
    u_before = ctl.get_u()
    step_now = ctl.step_info()["step"]
    na_g = ctl.get_act()
    e_g = ctl.get_inf()["e"]
    single.sample_only(step_now)
    e_s = single.get_inf()["e"]
    assert np.array_equal(e_g, e_s)          # eps depends on the global sample index only
    ref2 = oracle.step(p, x0, u_before, e_g, nthreads=4)
    assert np.allclose(na_g, ref2["next_act"], rtol=1e-5, atol=1e-6)
    # profiling mode must not deadlock a group (all shards are enqueued before any wait)
    ctl.set_profiling(True)
    ctl.get_act()
    kt = ctl.kernel_times()
    ctl.set_profiling(False)
    # peer-mailbox shards merge with ONE exchange (online softmax): no beta exchange, and the
    # exchange itself runs inside the last CTA of the averaging kernel -- no kernel of its own
    assert kt["comm_min"][1] == 0 and kt["comm_sum"][1] == 0 and kt["average"][1] == 1
    xt = ctl.exchange_times()
    assert xt["push_us"] > 0 and xt["merge_us"] > 0
    single.close()
    ctl.close()


@pytest.mark.gpu
def test_cpp_driver_on_a_device_group(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    subprocess.run(["make", "-C", os.path.join(ROOT, "cpp")], check=True, capture_output=True)
    r = subprocess.run([os.path.join(ROOT, "cpp", "mppi_main"), "-c",
                        os.path.join(ROOT, "config", "point_mass2d.yaml"), "--samples", "200000",
                        "--horizon", "100", "--steps", "100", "--quiet", "--devices", "0,1"],
                       capture_output=True, text=True, timeout=200)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Average controller execution time" in r.stdout
