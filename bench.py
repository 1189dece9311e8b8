#!/usr/bin/env python
"""bench.py -- rollout-steps/s of one MPPI control step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one full control step (sample -> rollout+cost -> beta/eta/weights -> weighted
average -> U update + shift -> next action), i.e. the reference's PointMassModel::get_act.
Default workload: point_mass3d, K=1e6 (GLOBAL, sharded over the N GPUs), T=200, A=3.

Printed JSON line (rank 0):
  value        K*T / (device time per step), graph-replayed steps back to back, inputs resident
               in HBM, CUDA events on the controller's stream, max over ranks
  e2e          the same metric through the reference-facing call sequence with HOST buffers:
               set_x(x_host) ; get_act(next_act_host) per step (blocking), wall clock
  roofline     the dominant kernel: algorithmic bytes / CUDA-event duration against
               MEASURED_PEAKS.json hbm_gbs.  With >= 4e5 samples on one GPU the whole step is ONE
               kernel (step_kernel: 4KTA written + 4KTA read back + 8K); otherwise the
               weighted-average kernel (part 4) of the kernel chain
  kernels      per-kernel average durations of the same chain (CUDA events between kernels,
               second timed region, direct launches instead of the graph)
  other_chains the same workload through the kernel chains (fused two-kernel, unfused
               three-kernel) with their per-kernel times and HBM fractions
  cpu_baseline the reference's own model/cost code (oracle/_ref) + oracle port of the
               reductions, 1 core, on a bounded sample of the workload
  parity_check one small K-sharded step of the SAME chain on the SAME ranks before the timing:
               noise and costs gathered from all ranks, single-shard oracle on rank 0 (checker
               only): S bit exact, beta / argmin exact, U within 1e-5, U bit-identical on every
               rank; a failure exits non-zero
  collectives_ms (N > 1) the NVLink exchange of the timed chain (push / wait-for-slowest /
               merge, %globaltimer stamps) AND, from a second controller on the same shards, the
               two NCCL all-reduces (min, sum) of the MPPI_COMM_NCCL chain
  configs      (N = 1) BASELINE.json configs[1] and [4] through the same API with host buffers:
               point_mass2d K=1e4 (ms_per_step, p50, p99) and the 1000-step closed loop at K=1e5
`--impl reference` times that CPU path with all host threads instead, on the full workload.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (K, T, A, dt, goal, w)  -- goal / cost.w from the reference's config/*.yaml
    "point_mass3d_K1e6_T200": (1000000, 200, 3, 0.1, [1, .5, .75, 0, 0, 0], [1, 1, 1, 5, 5, 5]),
    "point_mass2d_K1e4_T200": (10000, 200, 2, 0.1, [1, 0, 0, 0], [1, 1, 50, 50]),
    "point_mass1d_K1e4_T200": (10000, 200, 1, 0.1, [1, 0], [1, 5]),
    "point_mass2d_K1e5_T200": (100000, 200, 2, 0.1, [1, 0, 0, 0], [1, 1, 50, 50]),
    "point_mass3d_K1e7_T200": (10000000, 200, 3, 0.1, [1, .5, .75, 0, 0, 0], [1, 1, 1, 5, 5, 5]),
}
DEFAULT_WORKLOAD = "point_mass3d_K1e6_T200"
METRIC = "rollout_steps_per_s"
UNIT = "rollout-steps/s"


def workload_desc(name, K, T, A):
    return (f"{name}: K={K} global samples, T={T}, A={A}, S={2 * A}, dt=0.1, x0=0, U0=0, "
            f"lambda=1, sigma=0.025 (reference-compat preset), Philox seed 0")


def config_dict(name, K, T, A):
    """The `config` object of the JSON line -- identical for both arms (--impl ours|reference),
    so that the driver compares like with like; everything arm-specific goes to `details`."""
    eps_mb = 4.0 * K * T * A / 1e6
    return {"workload": workload_desc(name, K, T, A),
            "l2": ("inputs larger than L2: eps is %.0f MB per pass over all shards vs 126 MB of L2, "
                   "no flush between steps" % eps_mb) if eps_mb > 200 else
                  ("working set fits L2 (%.0f MB of eps): latency-bound config, no flush" % eps_mb)}


# --------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx = device_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.idx), "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, pw, reasons = [], None, [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # "under load" = samples in the upper half of the observed power range
        if sm:
            thr = (max(pw) + min(pw)) / 2 if pw else 0
            load = [c for c, p in zip(sm, pw) if p >= thr] or sm
            med = statistics.median(load)
        else:
            med = None
        return {"sm_mhz": med, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw) if pw else None}


# --------------------------------------------------------------------------------- CPU arm
def cpu_reference_step(po, K, T, A, dt, goal, w, x0, U, eps, nthreads):
    """One control step of the reference's CPU path on K samples; returns seconds."""
    t0 = time.perf_counter()
    if po.ref_available():
        S = po.ref_rollout_all(K, T, A, dt, 1.0, x0, U, goal, w, eps, nthreads=nthreads)
    else:
        p = po.make_problem(K, T, A, dt, goal, w)
        S = po.rollout_all(p, x0, U, eps, nthreads=nthreads)
    b, _ = po.beta(S)
    ex = po.exp(S, 1.0, b)
    eta32, _ = po.eta(ex)
    wts = po.weights(S, 1.0, b, eta32)
    un = po.update_act(U, wts, eps, K, T, A, nthreads=nthreads)
    po.shift(un, T, A)
    return time.perf_counter() - t0


def cpu_sample_inputs(K, T, A):
    eps = np.random.default_rng(0).standard_normal((K, T, A), dtype=np.float32)
    eps *= np.float32(0.025)
    return np.zeros(2 * A, np.float32), np.zeros((T, A), np.float32), eps


def run_reference(args, name, K, T, A, dt, goal, w):
    """The reference's own CPU implementation of the path on the box's host cores, all threads,
    on the FULL workload (same K, T, A as the product arm): at K=1e6 a step is ~1 s on 32
    threads.  Only if a single step would take longer than ~20 s (estimated from a 2 % probe)
    is K cut, and the line says so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as po
    po.lib()
    nthreads = os.cpu_count() or 1
    probe = max(1000, K // 50)
    x0, U, eps = cpu_sample_inputs(min(K, probe), T, A)
    cpu_reference_step(po, len(eps), T, A, dt, goal, w, x0, U, eps, nthreads)
    t_probe = cpu_reference_step(po, len(eps), T, A, dt, goal, w, x0, U, eps, nthreads)
    est = t_probe * K / len(eps)
    Ks = K if est <= 20.0 else max(probe, int(K * 20.0 / est))
    x0, U, eps = cpu_sample_inputs(Ks, T, A)
    for _ in range(args.warmup):
        cpu_reference_step(po, Ks, T, A, dt, goal, w, x0, U, eps, nthreads)
    times = [cpu_reference_step(po, Ks, T, A, dt, goal, w, x0, U, eps, nthreads)
             for _ in range(args.steps)]
    sec = sum(times) / len(times)
    val = Ks * T / sec
    kind = "reference" if po.ref_available() else "port"
    sample = (("the full workload, " if Ks == K else f"{Ks} of {K} samples per step (same T, A), ")
              + "rollout+cost = "
              + ("the reference's point_mass_gpu.cu+cost.cu compiled for the host (oracle/_ref), "
                 "OpenMP over samples" if kind == "reference" else "oracle port, OpenMP over samples")
              + "; beta/exp/eta/weights/shift = oracle port (serial), update_act_cpu = oracle port "
                "split over column ranges (bit-identical to the serial loop)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config_dict(name, K, T, A),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": nthreads, "kind": kind,
                         "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# --------------------------------------------------------------------------------- parity
PARITY_SHAPE = (20003, 40, 3, 5.0)      # K, T, A, lambda  (tests/dist_worker.py: gpu_mode)


def parity_check(m, capi, dist, torch, rank, world, local_rank, flags, comm):
    """One small step of the chain that is about to be timed, on the same ranks: the sampled
    noise and the costs of every shard are gathered, rank 0 runs the single-shard oracle on
    them (the checker; reference semantics: PointMassModel::get_act, src/point_mass.cu:129-203)
    and compares.  Bars: S bit exact, beta and argmin exact, eta / U / next action within 1e-5,
    replicated U bit-identical on every rank.  Returns the dict for the JSON line."""
    K, T, A, lam = PARITY_SHAPE
    goal, w = [1, .5, .75, 0, 0, 0], [1, 1, 1, 5, 5, 5]
    rs = np.random.RandomState(3)
    U = (0.1 * rs.standard_normal((T, A))).astype(np.float32)
    x0 = (0.05 * rs.standard_normal(2 * A)).astype(np.float32)
    if world > 1:
        from mppi_gpu_b200.torch_dist import sharded_controller
        ctl = sharded_controller(K, T, 0.1, 2 * A, A, comm=comm, device=local_rank, lam=lam,
                                 seed=11, flags=flags)
    else:
        ctl = m.PointMassModel(K, T, 0.1, 2 * A, A, lam=lam, seed=11, flags=flags, device=local_rank)
    ctl.memcpy_set_data(x0, U, goal, w)
    k0, k1 = capi.shard_range(K, rank, world)
    res = {"n": world, "shape": f"K={K},T={T},A={A},lambda={lam}", "flags": ctl.flags(), "steps": 2}
    ok, why = True, ""
    for step in range(2):
        pre = ctl.get_u()
        na = ctl.get_act()
        inf = ctl.get_inf()
        info = ctl.step_info()
        if world > 1:
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
            u_all = [torch.zeros(T * A + A, dtype=torch.float32, device="cuda") for _ in range(world)]
            mine = torch.from_numpy(np.concatenate([inf["u"].ravel(), na])).cuda()
            dist.all_gather(u_all, mine)
            same = all(torch.equal(u.view(torch.int32), u_all[0].view(torch.int32)) for u in u_all[1:])
        else:
            eps, cost, same = inf["e"], inf["cost"], True
        if rank == 0:
            from oracle import pyoracle as po                       # the checker, never the product
            po.lib()
            p = po.make_problem(K, T, A, 0.1, goal, w, lam=lam, arith=po.ARITH_FMA)
            ref = po.step(p, x0, pre, eps, nthreads=min(8, os.cpu_count() or 1))
            checks = {
                "S_bit_exact": bool(np.array_equal(cost.view(np.uint32), ref["S"].view(np.uint32))),
                "beta_exact": bool(np.float32(inf["beta"]).view(np.uint32) == np.float32(ref["beta"]).view(np.uint32)),
                "argmin_exact": bool(info["argmin"] == ref["argmin"]),
                "eta_1e-5": bool(abs(float(inf["nabla"]) - float(ref["eta"])) <= 1e-5 * float(ref["eta"])),
                "U_1e-5": bool(np.allclose(inf["u"], ref["U"], rtol=1e-5, atol=1e-6)),
                "next_act_1e-5": bool(np.allclose(na, ref["next_act"], rtol=1e-5, atol=1e-6)),
                "U_bit_identical_across_ranks": bool(same),
            }
            res["max_abs_dU"] = float(np.abs(inf["u"] - ref["U"]).max())
            bad = [k for k, v in checks.items() if not v]
            if bad:
                ok, why = False, f"step {step}: " + ",".join(bad)
            res["checks"] = checks
    ctl.close()
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
    if dist is not None:
        dist.broadcast(flag, 0)
    res["ok"] = bool(flag.item())
    if why:
        res["failed"] = why
    return res


# --------------------------------------------------------------------------------- extra configs
def closed_loop_config(m, capi, local_rank, nsteps=1000):
    """BASELINE.json configs[4]: receding-horizon point_mass2d, K=1e5, T=200, one CUDA graph per
    control step, the reference's loop (src/main.cu:326-371: get_act -> plant -> set_x) with the
    ideal double-integrator plant (the controller's own model) on the host."""
    K, T, A, dt, goal, w = WORKLOADS["point_mass2d_K1e5_T200"]
    ctl = m.PointMassModel(K, T, dt, 2 * A, A, seed=0, flags=capi.FLAG_AUTO_CHAIN, device=local_rank)
    x = np.zeros(2 * A, np.float32)
    ctl.memcpy_set_data(x, np.zeros((T, A), np.float32), goal, w)
    act = np.zeros(A, np.float32)
    for _ in range(20):
        ctl.get_act(act)
    ctl.memcpy_set_data(x, np.zeros((T, A), np.float32), goal, w)
    lat, plant = [], []
    for _ in range(nsteps):
        t0 = time.perf_counter()
        ctl.get_act(act)
        t1 = time.perf_counter()
        a = np.clip(act, -1.0, 1.0)
        x[:A] += dt * x[A:] + 0.5 * dt * dt * a
        x[A:] += dt * a
        ctl.set_x(x)
        t2 = time.perf_counter()
        lat.append(t1 - t0)
        plant.append(t2 - t1)
    flags = ctl.flags()
    ctl.close()
    lat = sorted(1e3 * v for v in lat)
    return {"workload": "point_mass2d, K=1e5, T=200, 1000 closed-loop control steps, graph per step",
            "p50_ms": lat[len(lat) // 2], "p99_ms": lat[int(0.99 * len(lat))], "max_ms": lat[-1],
            "plant_us": 1e6 * statistics.median(plant), "flags": flags,
            "final_state": [float(v) for v in x],
            "rollout_steps_per_s_p50": K * T / (lat[len(lat) // 2] * 1e-3)}


def reference_gpu_path_ms(K, T, A, dt, goal, w, nsteps=6):
    """BASELINE.json configs[1] asks for 'single B200 vs the reference sm_70-style GPU path':
    oracle/_ref/ref_gpu_run is the reference's own point_mass.cu + point_mass_gpu.cu + cost.cu +
    mppi_utils.cu, unmodified, recompiled -O3 for sm_100 behind a replay of src/main.cu:311-371
    (oracle/Makefile ref-gpu; built where /root/reference exists, travels as a binary).  It times
    its own get_act with a host clock as main.cu:329-332 does.  Returns the median ms per
    get_act after the first (None when the binary is not there)."""
    import struct
    import tempfile
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_gpu_run")
    if not os.path.exists(exe):
        return None
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        with open(fin, "wb") as f:
            f.write(struct.pack("<4i", K, T, A, nsteps))
            f.write(struct.pack("<f", dt))
            for arr in (np.zeros(2 * A), np.zeros(T * A), goal, w):
                f.write(np.asarray(arr, np.float32).tobytes())
        try:
            r = subprocess.run([exe, fin, fout], capture_output=True, text=True, timeout=120)
            if r.returncode != 0:
                return None
            raw = np.fromfile(fout, np.float32)
        except Exception:
            return None
    return float(np.median(raw[-nsteps:][1:]))


def small_config(m, capi, local_rank, name, nsteps=200):
    """BASELINE.json configs[1] (and [0]'s shape on the GPU): device time per step and the
    latency of set_x + get_act with host buffers."""
    K, T, A, dt, goal, w = WORKLOADS[name]
    ctl = m.PointMassModel(K, T, dt, 2 * A, A, seed=0, flags=capi.FLAG_AUTO_CHAIN, device=local_rank)
    x = np.zeros(2 * A, np.float32)
    ctl.memcpy_set_data(x, np.zeros((T, A), np.float32), goal, w)
    act = np.zeros(A, np.float32)
    for _ in range(20):
        ctl.get_act(act)
    ctl.timer_start()
    for _ in range(nsteps):
        ctl.step_enqueue()
    ms = ctl.timer_stop() / nsteps
    ctl.step_wait()
    lat = []
    for _ in range(nsteps):
        t0 = time.perf_counter()
        ctl.set_x(x)
        ctl.get_act(act)
        lat.append(time.perf_counter() - t0)
        x[:A] = 1e-3 * act
    flags = ctl.flags()
    ctl.close()
    lat = sorted(1e3 * v for v in lat)
    out = {"ms_per_step": ms, "rollout_steps_per_s": K * T / (ms * 1e-3), "p50_ms": lat[len(lat) // 2],
           "p99_ms": lat[int(0.99 * len(lat))], "flags": flags}
    if A == 2:      # the only action dim whose average is not defective in the reference
        ref_ms = reference_gpu_path_ms(K, T, A, dt, goal, w)
        if ref_ms:
            out["reference_gpu_path"] = {
                "what": "the reference's own GPU path recompiled for sm_100 (oracle/_ref/ref_gpu_run), "
                        "same GPU, same shape, its own host clock around get_act",
                "ms_per_get_act": ref_ms, "rollout_steps_per_s": K * T / (ref_ms * 1e-3),
                "speedup_p50": ref_ms / lat[len(lat) // 2]}
    return out


# --------------------------------------------------------------------------------- ours
def run_ours(args, name, K, T, A, dt, goal, w):
    import torch
    import mppi_gpu_b200 as m
    from mppi_gpu_b200 import capi

    os.environ.setdefault("NCCL_DEBUG", "WARN")     # stdout carries the JSON line and nothing else
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # chain selection: MPPI_FLAG_AUTO_CHAIN lets the library choose from the shard's shape; the
    # choice is read back and reported (details.flags / details.chain)
    flags = args.flags if args.flags >= 0 else capi.FLAG_AUTO_CHAIN
    if world > 1:
        from mppi_gpu_b200.torch_dist import sharded_controller
        ctl = sharded_controller(K, T, dt, 2 * A, A, comm=args.comm, device=local_rank, seed=0,
                                 flags=flags)
    else:
        ctl = m.PointMassModel(K, T, dt, 2 * A, A, seed=0, flags=flags, device=local_rank)
    flags = ctl.flags()                       # MPPI_FLAG_AUTO_CHAIN resolved by the library
    x0 = np.zeros(2 * A, np.float32)
    ctl.memcpy_set_data(x0, np.zeros((T, A), np.float32), goal, w)

    # ---- parity first: the chain about to be timed, on these ranks, against the oracle
    parity = None
    if not args.no_parity_check:
        parity = parity_check(m, capi, dist, torch, rank, world, local_rank, flags, args.comm)
        if not parity["ok"]:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": world,
                                  "parity_check": parity, "error": "parity check failed"}))
            if dist is not None:
                dist.destroy_process_group()
            raise SystemExit(3)

    # ---- warm-up (also instantiates the CUDA graph)
    for _ in range(max(args.warmup, 3)):
        ctl.get_act()
    barrier()

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)

    # ---- region 1: device-resident steps, graph replays back to back
    launches0 = ctl.launch_count()
    barrier()
    ctl.timer_start()
    for _ in range(args.steps):
        ctl.step_enqueue()
    ms_total = ctl.timer_stop()
    ctl.step_wait()
    launches = ctl.launch_count() - launches0
    barrier()
    ms_step = max_over_ranks(ms_total / args.steps)
    value = K * T / (ms_step * 1e-3)

    # ---- region 2: end to end through the reference-facing calls with host buffers
    x_host = np.zeros(2 * A, np.float32)
    act_host = np.zeros(A, np.float32)
    lat = []
    xt = {"push_us": [], "wait_slowest_us": [], "merge_us": []}
    barrier()
    t_begin = time.perf_counter()
    for i in range(args.steps):
        t0 = time.perf_counter()
        ctl.set_x(x_host)                 # H2D, S floats (reference: set_x, src/main.cu:371)
        ctl.get_act(act_host)             # blocking; D2H of A floats (reference: get_act)
        lat.append(time.perf_counter() - t0)
        x_host[:A] = 1e-3 * act_host      # host-side dependency on the result
    e2e_sec = (time.perf_counter() - t_begin) / args.steps
    barrier()
    e2e_sec = max_over_ranks(e2e_sec)
    e2e_val = K * T / e2e_sec
    lat_ms = sorted(1e3 * x for x in lat)
    if world > 1 and args.comm == "p2p":
        # phases of the in-kernel NVLink exchange (%globaltimer stamps of the last step)
        for _ in range(min(args.steps, 20)):
            ctl.get_act(act_host)
            for k, v in ctl.exchange_times().items():
                xt[k].append(v)

    # ---- region 3: the same chain with CUDA events between the kernels
    ctl.set_profiling(True)
    for _ in range(args.steps):
        ctl.get_act()
    kt = ctl.kernel_times()
    ctl.set_profiling(False)
    kernels = {k: (ms / n) for k, (ms, n) in kt.items() if n}
    clk = clocks.stop() if rank == 0 else None

    k_local = ctl.k_local
    tile_kernel = bool(flags & capi.FLAG_TILE_KERNEL) and launches == args.steps
    one_kernel = tile_kernel or (bool(flags & capi.FLAG_STEP_KERNEL) and launches == args.steps)
    avg_ms = max_over_ranks(kernels["average"])
    units = float(k_local) * T                      # rollout-steps one launch of this shard processes
    if one_kernel:
        # The step IS this one kernel, so its average launch duration is taken from region 1
        # (K launches back to back between two CUDA events) rather than from region 3, whose
        # per-launch event pairs add the launch gap.
        avg_ms = ms_step
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # Algorithmic bytes.  SURVEY.md 8(d): 12*A bytes per rollout-step (+16/T) -- eps written by the
    # sampler, read by the rollout, read by the average.  A chain that hands eps from the sampler
    # to the rollout in registers moves 8*A (+8/T): round 1's accounting, kept as `frac_fused_design`.
    # The dominant kernel of the unfused chains (average_kernel) is accounted with its own 4*A.
    survey_bytes = 12.0 * A * units + 16.0 * k_local
    fused_bytes = 8.0 * A * units + 8.0 * k_local
    if one_kernel:
        alg_bytes = survey_bytes
    else:
        alg_bytes = 4.0 * A * units + 4.0 * k_local
    achieved = alg_bytes / (avg_ms * 1e-3) / 1e9
    traffic_file = ("tile_traffic.json" if tile_kernel else "step_traffic.json" if one_kernel
                    else "average_traffic.json")
    ncu_traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", traffic_file)))
        if world == 1 and name == DEFAULT_WORKLOAD:
            ncu_traffic = tj.get("dram_bytes_per_launch")
            traffic_src = {k: tj.get(k) for k in ("source", "commit", "kernel", "captured") if k in tj}
            traffic_src["file"] = "profiles/" + traffic_file
    except Exception:
        pass
    if tile_kernel:
        kname = ("tile_kernel (parts 1-5 in one persistent kernel; eps drawn into shared memory, "
                 "integrated and averaged there -- it never reaches HBM)")
        note = ("achieved = SURVEY 8(d) algorithmic bytes (12*A per rollout-step) / launch time: the "
                "kernel beats the HBM roofline of the path (frac > 1) because it does not move those "
                "bytes -- measured DRAM traffic in `traffic`; what bounds it is instruction issue "
                "(Philox + Box-Muller), see `issue`")
    elif one_kernel:
        kname = ("step_kernel (parts 1-5 in one persistent kernel: eps written by the rollout "
                 "warps, read back by the TMA-fed average warps)")
        note = "frac follows SURVEY 8(d) (12*A); the kernel itself moves 8*A: frac_fused_design"
    else:
        kname, note = "average_kernel (part 4: sum_k w_k eps_k[t,a])", "4*A bytes per rollout-step: this pass only"
    roofline = {"kernel": kname, "bound": "hbm",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback",
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": avg_ms,
                "traffic": ncu_traffic, "traffic_source": traffic_src, "note": note}
    if one_kernel:
        roofline["frac_survey_bytes"] = survey_bytes / (avg_ms * 1e-3) / 1e9 / peak
        roofline["frac_fused_design"] = fused_bytes / (avg_ms * 1e-3) / 1e9 / peak
        roofline["survey_bytes_per_launch"] = survey_bytes
        roofline["fused_design_bytes_per_launch"] = fused_bytes
    if tile_kernel:
        try:
            roofline["issue"] = json.load(open(os.path.join(ROOT, "profiles", "tile_issue.json")))
        except Exception:
            roofline["issue"] = None
    eps_bytes = 4.0 * k_local * T * A
    # The sampling work is bound by the FMA pipe of the SM sub-partitions, not by HBM
    # (tools/ubench/pipes.cu): per 128-sample warp and time step, A Philox calls of
    # 16.33 IMAD.WIDE (4 pipe cycles each; 20 per call less the loop-invariant first rounds) + 12
    # scalar FP, and for the rollout 24A+4 packed FP32x2 operations (2 cycles each).
    sm_clk = ((clk or {}).get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0) * 1e6
    n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count
    warps = math.ceil(k_local / 128.0)

    def fma_pipe(ms, rollout, sampling):
        cyc = warps * T * ((16.33 * 4 + 12) * A * (1 if sampling else 0) + (48.0 * A + 8) * (1 if rollout else 0))
        return {"algorithmic_cycles_per_subpartition": cyc / (4 * n_sm),
                "frac": cyc / (4 * n_sm) / (ms * 1e-3 * sm_clk)}

    per_kernel = {}
    for kname_, ms in kernels.items():
        d = {"ms": ms}
        if kname_ in ("sample", "rollout"):
            d["hbm_gbs"] = eps_bytes / (ms * 1e-3) / 1e9
            d["hbm_frac"] = d["hbm_gbs"] / peak
        if kname_ == "sample":
            d["fma_pipe"] = fma_pipe(ms, False, True)
        if kname_ == "rollout" and (flags & capi.FLAG_FUSED_SAMPLING):
            d["fma_pipe"] = fma_pipe(ms, True, True)
        per_kernel[kname_] = d

    chain = ("one kernel, eps on chip: generator warps -> shared-memory tile -> integrator warp -> "
             "averaging threads -> merge+finalize" + (" + NVLink exchange in the last CTA" if world > 1 else "")
             if tile_kernel else
             "one kernel: sample+rollout warps || weights+average warps -> merge+finalize"
             + (" + NVLink exchange in the last CTA" if world > 1 else "")
             if one_kernel else
             "sample+rollout(fused) -> weights -> average -> finalize"
             if flags & capi.FLAG_FUSED_SAMPLING else
             "rollout -> weights -> average -> finalize; the sampler of step n+1 runs "
             "behind step n's chain on a second stream (pipelined sampling)"
             if flags & capi.FLAG_PIPELINED_SAMPLING else
             "sample -> rollout -> weights -> average -> finalize")
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(name, K, T, A),
        "details": {"k_local": k_local, "flags": flags, "graph": not (flags & capi.FLAG_NO_GRAPH),
                    "chain": chain,
                    "timing": "value: CUDA events on the controller stream around K graph launches; "
                              "kernels/roofline: CUDA events between kernels in a second region of K steps"},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": 4 * 2 * A,
                "d2h_bytes_per_step": 4 * A, "ms_per_step": e2e_sec * 1e3},
        "latency_ms": {"p50": lat_ms[len(lat_ms) // 2], "p99": lat_ms[min(len(lat_ms) - 1, int(0.99 * len(lat_ms)))],
                       "min": lat_ms[0], "max": lat_ms[-1]},
        "gpu_launches": launches,
        "parity_check": parity,
        "roofline": roofline,
        "kernels": per_kernel,
        "clocks": clk,
    }
    if one_kernel:
        per_kernel = {("tile" if tile_kernel else "step"): {
            "ms": avg_ms, "alg_gbs": achieved, "alg_frac": achieved / peak,
            "ms_with_event_pairs": kernels["average"]}}
        out["kernels"] = per_kernel
    if world == 1 and name == DEFAULT_WORKLOAD and not args.no_other_chains:
        # the other chains timed beside it on the same workload: the HBM one-kernel step, the
        # fused two-kernel chain and the canonical unfused one (sample, rollout, average)
        ctl.close()
        ctl = None
        base = flags & ~(capi.FLAG_FUSED_SAMPLING | capi.FLAG_STEP_KERNEL | capi.FLAG_TILE_KERNEL |
                         capi.FLAG_PIPELINED_SAMPLING)
        others = {}
        for cname, cflags in (("tile_kernel_eps_on_chip", base | capi.FLAG_TILE_KERNEL),
                              ("step_kernel_eps_via_hbm", base | capi.FLAG_STEP_KERNEL),
                              ("fused_2_kernels", base | capi.FLAG_FUSED_SAMPLING),
                              ("unfused_3_kernels", base),
                              # an OPTION, not the headline: Philox-4x32-7 (mppi_params.philox_rounds)
                              ("fused_2_kernels_philox7", base | capi.FLAG_FUSED_SAMPLING)):
            if cflags == flags and not cname.endswith("philox7"):
                continue
            c2 = m.PointMassModel(K, T, dt, 2 * A, A, seed=0, flags=cflags, device=local_rank,
                                  philox_rounds=7 if cname.endswith("philox7") else 10)
            c2.memcpy_set_data(x0, np.zeros((T, A), np.float32), goal, w)
            for _ in range(3):
                c2.get_act()
            c2.timer_start()
            for _ in range(args.steps):
                c2.step_enqueue()
            ms4 = c2.timer_stop() / args.steps
            c2.step_wait()
            c2.set_profiling(True)
            for _ in range(args.steps):
                c2.get_act()
            kt4 = {k: ms / n for k, (ms, n) in c2.kernel_times().items() if n}
            c2.set_profiling(False)
            c2.close()
            one = bool(cflags & (capi.FLAG_STEP_KERNEL | capi.FLAG_TILE_KERNEL))
            others[cname] = {
                "ms_per_step": ms4, "value": K * T / (ms4 * 1e-3),
                "frac_survey_bytes": survey_bytes / (ms4 * 1e-3) / 1e9 / peak,
                "kernels": {k: ({"ms": v, "hbm_gbs": eps_bytes / (v * 1e-3) / 1e9,
                                 "hbm_frac": eps_bytes / (v * 1e-3) / 1e9 / peak}
                                if (k in ("sample", "rollout", "average") and not one) else {"ms": v})
                            for k, v in kt4.items()}}
            if not one:
                if "sample" in kt4:
                    others[cname]["kernels"]["sample"]["fma_pipe"] = fma_pipe(kt4["sample"], False, True)
                if "rollout" in kt4 and (cflags & capi.FLAG_FUSED_SAMPLING) and not cname.endswith("philox7"):
                    others[cname]["kernels"]["rollout"]["fma_pipe"] = fma_pipe(kt4["rollout"], True, True)
        out["other_chains"] = others
    if world > 1:
        med = lambda v: statistics.median(v) if v else None
        coll = {"comm": args.comm}
        if args.comm == "p2p":
            coll["p2p"] = {
                "kind": "ONE exchange per step over NVLink peer mailboxes (direct P2P stores + flags), "
                        + ("inside the last CTA of the step's kernel" if one_kernel else
                           "inside the last CTA of average_kernel")
                        + ": every shard averages relative to its own minimum, the exchange rescales "
                          "by exp(-(beta_r-beta)/lambda), sums in rank order and applies the U update",
                "push_us": med(xt["push_us"]), "wait_slowest_us": med(xt["wait_slowest_us"]),
                "merge_us": med(xt["merge_us"]),
                "kernel_ms": kernels.get("comm_sum")}
        else:
            coll["nccl"] = {"min_u64": kernels.get("comm_min"), "sum_i64": kernels.get("comm_sum")}
        if args.comm == "p2p" and not args.no_nccl_leg:
            # the two NCCL all-reduces of the MPPI_COMM_NCCL chain, measured on a second
            # controller over the same shards (north_star: reported separately)
            if ctl is not None:
                ctl.close()
                ctl = None
            from mppi_gpu_b200.torch_dist import sharded_controller
            cn = sharded_controller(K, T, dt, 2 * A, A, comm="nccl", device=local_rank, seed=0,
                                    flags=capi.FLAG_FUSED_SAMPLING)
            cn.memcpy_set_data(x0, np.zeros((T, A), np.float32), goal, w)
            for _ in range(5):
                cn.get_act()
            barrier()
            cn.timer_start()
            for _ in range(args.steps):
                cn.step_enqueue()
            ms_n = cn.timer_stop() / args.steps
            cn.step_wait()
            barrier()
            cn.set_profiling(True)
            for _ in range(20):
                cn.get_act()
            ktn = {k: ms / n for k, (ms, n) in cn.kernel_times().items() if n}
            cn.set_profiling(False)
            cn.close()
            coll["nccl"] = {
                "kind": "ncclAllReduce(min, 1 x u64 packed key) + ncclAllReduce(sum, (T*A+1) x int64) "
                        "captured in the CUDA graph of the fused two-kernel chain",
                "min_u64": max_over_ranks(ktn.get("comm_min", 0.0)),
                "sum_i64": max_over_ranks(ktn.get("comm_sum", 0.0)),
                "ms_per_step": max_over_ranks(ms_n),
                "value": K * T / (max_over_ranks(ms_n) * 1e-3)}
        out["collectives_ms"] = coll
        out["details"]["comm"] = args.comm
        if not args.no_weak_probe and name == DEFAULT_WORKLOAD:
            # BASELINE.json fixes K globally (strong scaling: the shards shrink with N and fall below
            # the size at which the kernels run at their large-K efficiency).  Beside it, the same
            # ranks with the single-GPU shard size each (K = N x 1e6): what N GPUs sustain when the
            # problem is allowed to grow.  Extra information, not the headline value.
            if ctl is not None:
                ctl.close()
                ctl = None
            from mppi_gpu_b200.torch_dist import sharded_controller
            cw = sharded_controller(K * world, T, dt, 2 * A, A, comm=args.comm, device=local_rank, seed=0,
                                    flags=capi.FLAG_AUTO_CHAIN)
            cw.memcpy_set_data(x0, np.zeros((T, A), np.float32), goal, w)
            for _ in range(5):
                cw.get_act()
            barrier()
            cw.timer_start()
            for _ in range(args.steps):
                cw.step_enqueue()
            ms_w = cw.timer_stop() / args.steps
            cw.step_wait()
            barrier()
            ms_w = max_over_ranks(ms_w)
            out["details"]["weak_scaling_probe"] = {
                "K_global": K * world, "k_local": cw.k_local, "flags": cw.flags(), "ms_per_step": ms_w,
                "value": K * world * T / (ms_w * 1e-3), "unit": UNIT}
            cw.close()

    # ---- BASELINE.json configs[1] and [4] through the same API (N=1 only; < 2 s)
    if rank == 0 and world == 1 and name == DEFAULT_WORKLOAD and not args.no_extra_configs:
        if ctl is not None:
            ctl.close()
            ctl = None
        out["configs"] = {
            "point_mass1d_K1e4_T200": small_config(m, capi, local_rank, "point_mass1d_K1e4_T200"),
            "point_mass2d_K1e4_T200": small_config(m, capi, local_rank, "point_mass2d_K1e4_T200"),
            "point_mass2d_K1e5_T200_closed_loop_1000": closed_loop_config(m, capi, local_rank)}

    # ---- CPU baseline beside it (rank 0, N=1 only): 1 core, bounded sample
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import pyoracle as po
        po.lib()
        Ks = min(K, 50000)
        cx0, cU, ceps = cpu_sample_inputs(Ks, T, A)
        cpu_reference_step(po, min(Ks, 2000), T, A, dt, goal, w, cx0, cU, ceps[:2000], 1)
        ts = [cpu_reference_step(po, Ks, T, A, dt, goal, w, cx0, cU, ceps, 1) for _ in range(3)]
        sec = statistics.median(ts)
        kind = "reference" if po.ref_available() else "port"
        out["cpu_baseline"] = {
            "value": Ks * T / sec, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"{Ks} of {K} samples (same T, A), median of 3 serial steps, "
                      f"{os.cpu_count()} host cores present; rollout+cost = "
                      + ("reference sources compiled for the host (oracle/_ref)" if kind == "reference"
                         else "oracle port") + ", reductions/update = oracle port"}
    if ctl is not None:
        ctl.close()
    if rank == 0:
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--flags", type=int, default=-1,
                    help="MPPI_FLAG_* bits; default MPPI_FLAG_AUTO_CHAIN (the library picks the chain "
                         "from the shard's shape and reports it in details.flags)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--no-other-chains", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true")
    ap.add_argument("--no-nccl-leg", action="store_true")
    ap.add_argument("--no-weak-probe", action="store_true")
    ap.add_argument("--comm", default="p2p", choices=["p2p", "nccl"],
                    help="K-shard exchange for --gpus > 1: NVLink peer mailboxes or NCCL")
    args = ap.parse_args()
    K, T, A, dt, goal, w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, args.workload, K, T, A, dt, goal, w)
    else:
        run_ours(args, args.workload, K, T, A, dt, goal, w)


if __name__ == "__main__":
    main()
