"""Worker for the multi-rank tests (launched by torch.distributed.run; not a pytest file).

mode "gloo": CPU-only check of the K-shard exchange logic that the C ABI implements on the
  GPU: shard ranges from mppi_shard_range, per-shard oracle rollouts, all-reduce(min) of the
  packed (ordered cost, global index) key, all-reduce(sum) of the 2^30 fixed-point
  accumulators; the result must equal the single-shard oracle step.
mode "nccl" / "p2p": the real thing on GPUs -- every rank owns one shard of one controller,
  exchanging through NCCL all-reduces or through the NVLink peer mailboxes.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import REF_CFG, make_inputs  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

ACC_SCALE = 2.0 ** 30


def ordered(f):
    u = np.float32(f).view(np.uint32).astype(np.uint64)
    return np.where(u & 0x80000000, (~u) & 0xFFFFFFFF, u | 0x80000000).astype(np.uint64)


def gloo_mode():
    from mppi_gpu_b200 import capi
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    K, T, A, lam = 1003, 25, 2, 4.0
    cfg = REF_CFG[A]
    x0, U, eps = make_inputs(K, T, A, seed=77, sigma=0.3)
    k0, k1 = capi.shard_range(K, rank, world)
    p = po.make_problem(k1 - k0, T, A, 0.1, cfg["goal"], cfg["w"], lam=lam)
    S = po.rollout_all(p, x0, U, eps[k0:k1])
    # (3) beta: all-reduce(min) of the packed key
    keys = (ordered(S) << np.uint64(32)) | np.arange(k0, k1, dtype=np.uint64)
    key = torch.tensor([int(keys.min()) - (1 << 63)], dtype=torch.int64)   # order-preserving bias
    dist.all_reduce(key, op=dist.ReduceOp.MIN)
    gkey = int(key.item()) + (1 << 63)
    o = np.uint32(gkey >> 32)
    beta = (np.uint32(o & 0x7FFFFFFF) if o & 0x80000000 else np.uint32(~o)).view(np.float32)
    argmin = gkey & 0xFFFFFFFF
    # (4) weighted partial sums + eta in fixed point, all-reduce(sum) int64
    wt = po.exp(S, lam, beta)
    num = (wt[:, None].astype(np.float64) * eps[k0:k1].reshape(k1 - k0, T * A)).sum(0)
    acc = np.rint(np.concatenate([num.astype(np.float32), [wt.sum(dtype=np.float32)]])
                  .astype(np.float64) * ACC_SCALE).astype(np.int64)
    t = torch.from_numpy(acc)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    tot = t.numpy().astype(np.float64) / ACC_SCALE
    Un = U.ravel() + (tot[:-1] / tot[-1]).astype(np.float32)
    Un = po.shift(Un, T, A)
    # single-shard oracle
    pf = po.make_problem(K, T, A, 0.1, cfg["goal"], cfg["w"], lam=lam)
    ref = po.step(pf, x0, U, eps)
    assert argmin == ref["argmin"], (argmin, ref["argmin"])
    assert beta == ref["beta"]
    assert np.allclose(Un, ref["U"].ravel(), rtol=1e-5, atol=1e-6), np.abs(Un - ref["U"].ravel()).max()
    assert abs(tot[-1] - float(ref["eta"])) <= 1e-5 * float(ref["eta"])
    dist.barrier()
    if rank == 0:
        print("GLOO_OK")
    dist.destroy_process_group()


def gloo_merge_mode():
    """CPU mirror of xchg_merge_body (csrc/xchg.cuh): shard-local softmax reference,
    all-gather of {key, fixed-point accumulators}, rescale, sum in rank order."""
    from mppi_gpu_b200 import capi
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    K, T, A, lam = 1501, 19, 3, 0.7
    cfg = REF_CFG[A]
    x0, U, eps = make_inputs(K, T, A, seed=78, sigma=0.3)
    k0, k1 = capi.shard_range(K, rank, world)
    p = po.make_problem(k1 - k0, T, A, 0.1, cfg["goal"], cfg["w"], lam=lam)
    S = po.rollout_all(p, x0, U, eps[k0:k1])
    keys = (ordered(S) << np.uint64(32)) | np.arange(k0, k1, dtype=np.uint64)
    my_key = int(keys.min())
    beta_r = S.min()
    wt = po.exp(S, lam, beta_r)                       # relative to the shard's own minimum
    num = (wt[:, None].astype(np.float64) * eps[k0:k1].reshape(k1 - k0, T * A)).sum(0)
    acc = np.rint(np.concatenate([num.astype(np.float32), [wt.sum(dtype=np.float32)]])
                  .astype(np.float64) * ACC_SCALE).astype(np.int64)
    msg = torch.from_numpy(np.concatenate([[my_key - (1 << 63)], acc]).astype(np.int64))
    box = [torch.zeros_like(msg) for _ in range(world)]
    dist.all_gather(box, msg)
    gkeys = [int(b[0].item()) + (1 << 63) for b in box]
    gkey = min(gkeys)

    def key_beta(k):
        o = np.uint32(k >> 32)
        return (np.uint32(o & 0x7FFFFFFF) if o & 0x80000000 else np.uint32(~o)).view(np.float32)

    beta = key_beta(gkey)
    nil = np.float32(-(np.float32(1) / np.float32(lam)))
    tot = np.zeros(T * A + 1, np.float64)
    for r in range(world):
        f = np.exp(np.float32(nil * np.float32(key_beta(gkeys[r]) - beta)), dtype=np.float32)
        tot += box[r][1:].numpy().astype(np.float64) * float(f)
    tot = np.rint(tot) / ACC_SCALE
    Un = U.ravel() + (tot[:-1] / tot[-1]).astype(np.float32)
    Un = po.shift(Un, T, A)
    pf = po.make_problem(K, T, A, 0.1, cfg["goal"], cfg["w"], lam=lam)
    ref = po.step(pf, x0, U, eps)
    assert (gkey & 0xFFFFFFFF) == ref["argmin"] and beta == ref["beta"]
    assert np.allclose(Un, ref["U"].ravel(), rtol=1e-5, atol=1e-6), np.abs(Un - ref["U"].ravel()).max()
    assert abs(tot[-1] - float(ref["eta"])) <= 1e-5 * float(ref["eta"])
    dist.barrier()
    if rank == 0:
        print("GLOO_MERGE_OK")
    dist.destroy_process_group()


def gpu_mode(comm, flags=0):
    from mppi_gpu_b200 import capi
    from mppi_gpu_b200.torch_dist import sharded_controller
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    K, T, A, lam = 20003, 40, 3, 5.0
    cfg = REF_CFG[A]
    x0, U, _ = make_inputs(K, T, A, seed=3)
    ctl = sharded_controller(K, T, 0.1, 2 * A, A, comm=comm, device=local, lam=lam, seed=11,
                             flags=flags)
    k0, k1 = capi.shard_range(K, rank, world)
    assert (ctl.k_offset, ctl.k_local) == (k0, k1 - k0)
    ctl.memcpy_set_data(x0, U, cfg["goal"], cfg["w"])
    p = po.make_problem(K, T, A, 0.1, cfg["goal"], cfg["w"], lam=lam, arith=po.ARITH_FMA)
    for step in range(3):
        Uo = ctl.get_u()                       # the (replicated) U this step starts from
        na = ctl.get_act()
        inf = ctl.get_inf()
        info = ctl.step_info()
        # the sampled noise depends on the GLOBAL sample index only
        want = po.sample_eps(11, step, k0, k1 - k0, T, A, [0.025] * A)
        d = np.abs(inf["e"] - want) / 0.025
        assert d.max() < 1e-3 and np.mean(d > 2e-5) < 1e-5, (d.max(), np.mean(d > 2e-5))
        # gather every shard's noise and costs on all ranks (NCCL all_gather needs equal
        # sizes: pad to the largest shard), run the single-shard oracle
        sizes = [capi.shard_range(K, r, world)[1] - capi.shard_range(K, r, world)[0]
                 for r in range(world)]
        kmax = max(sizes)
        e_pad = torch.zeros(kmax * T * A, dtype=torch.float32, device="cuda")
        c_pad = torch.zeros(kmax, dtype=torch.float32, device="cuda")
        e_pad[: (k1 - k0) * T * A] = torch.from_numpy(inf["e"].ravel()).cuda()
        c_pad[: k1 - k0] = torch.from_numpy(inf["cost"]).cuda()
        e_all = [torch.zeros_like(e_pad) for _ in range(world)]
        c_all = [torch.zeros_like(c_pad) for _ in range(world)]
        dist.all_gather(e_all, e_pad)
        dist.all_gather(c_all, c_pad)
        eps = torch.cat([e[: n * T * A] for e, n in zip(e_all, sizes)]).cpu().numpy().reshape(K, T, A)
        cost = torch.cat([c[:n] for c, n in zip(c_all, sizes)]).cpu().numpy()
        ref = po.step(p, x0, Uo, eps, nthreads=4)
        assert np.array_equal(cost.view(np.uint32), ref["S"].view(np.uint32))
        assert info["argmin"] == ref["argmin"] and inf["beta"] == ref["beta"]
        assert abs(float(inf["nabla"]) - float(ref["eta"])) <= 1e-5 * float(ref["eta"])
        assert np.allclose(inf["u"], ref["U"], rtol=1e-5, atol=1e-6)
        assert np.allclose(na, ref["next_act"], rtol=1e-5, atol=1e-6)
        # replicated state is bit-identical on every rank
        u_all = [torch.zeros(T * A, dtype=torch.float32, device="cuda") for _ in range(world)]
        dist.all_gather(u_all, torch.from_numpy(inf["u"].ravel()).cuda())
        for u in u_all[1:]:
            assert torch.equal(u, u_all[0])
    ctl.close()
    dist.barrier()
    if rank == 0:
        print(comm.upper() + "_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    import signal
    import traceback
    signal.alarm(150)                 # never outlive the test: a dead peer must not hang us
    try:
        if sys.argv[1] == "gloo":
            gloo_mode()
        elif sys.argv[1] == "gloo_merge":
            gloo_merge_mode()
        else:
            gpu_mode(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    except BaseException:
        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)
