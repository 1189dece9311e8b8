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
`--impl reference` times that CPU path with all host threads instead.
"""
import argparse
import json
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
    rs = np.random.RandomState(0)
    eps = (0.025 * rs.standard_normal((K, T, A))).astype(np.float32)
    return np.zeros(2 * A, np.float32), np.zeros((T, A), np.float32), eps


def run_reference(args, name, K, T, A, dt, goal, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as po
    po.lib()
    nthreads = os.cpu_count() or 1
    Ks = min(K, 100000)
    x0, U, eps = cpu_sample_inputs(Ks, T, A)
    for _ in range(args.warmup):
        cpu_reference_step(po, Ks, T, A, dt, goal, w, x0, U, eps, nthreads)
    times = [cpu_reference_step(po, Ks, T, A, dt, goal, w, x0, U, eps, nthreads)
             for _ in range(args.steps)]
    sec = sum(times) / len(times)
    val = Ks * T / sec
    kind = "reference" if po.ref_available() else "port"
    sample = (f"{Ks} of {K} samples per step (same T, A); rollout+cost = "
              + ("the reference's point_mass_gpu.cu+cost.cu compiled for the host (oracle/_ref), "
                 "OpenMP over samples" if kind == "reference" else "oracle port, OpenMP over samples")
              + "; beta/exp/eta/weights/shift = oracle port (serial), update_act_cpu = oracle port "
                "split over column ranges (bit-identical to the serial loop)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": workload_desc(name, K, T, A)},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": nthreads, "kind": kind,
                         "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# --------------------------------------------------------------------------------- ours
def run_ours(args, name, K, T, A, dt, goal, w):
    import torch
    import mppi_gpu_b200 as m
    from mppi_gpu_b200 import capi

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

    # chain selection: with >= 4e5 samples per GPU the sampling is fused into the rollout (one
    # pass writes eps and integrates; identical eps values) -- on a single shard as the
    # one-kernel step, where the weighted average of finished tiles overlaps the rollout of
    # the next ones; below that the step is latency-bound and the unfused chain with the TMA
    # rollout wins.
    k_loc = capi.shard_range(K, rank, world)
    k_loc = k_loc[1] - k_loc[0]
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

    # ---- region 3: the same chain with CUDA events between the kernels
    ctl.set_profiling(True)
    for _ in range(args.steps):
        ctl.get_act()
    kt = ctl.kernel_times()
    ctl.set_profiling(False)
    kernels = {k: (ms / n) for k, (ms, n) in kt.items() if n}
    clk = clocks.stop() if rank == 0 else None

    k_local = ctl.k_local
    one_kernel = bool(flags & capi.FLAG_STEP_KERNEL) and launches == args.steps * (1 if world == 1 else 2)
    avg_ms = max_over_ranks(kernels["average"])
    if one_kernel:
        # eps written once and read back once, S written and read once.  The step IS this one
        # kernel, so its average launch duration is taken from region 1 (K launches back to
        # back between two CUDA events; includes the 20-byte D2H node) rather than from region
        # 3, whose per-launch event pairs add the launch gap.
        alg_bytes = 8.0 * k_local * T * A + 8.0 * k_local
        if world == 1:
            avg_ms = ms_step
    else:
        alg_bytes = 4.0 * k_local * T * A + 4.0 * k_local
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (avg_ms * 1e-3) / 1e9
    ncu_traffic = None
    try:
        ncu_traffic = json.load(open(os.path.join(
            ROOT, "profiles", "step_traffic.json" if one_kernel else "average_traffic.json"))).get(
            "dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"kernel": ("step_kernel (parts 1-5 in one persistent kernel: eps written by the rollout "
                           "warps, read back by the TMA-fed average warps)" if one_kernel else
                           "average_kernel (part 4: sum_k w_k eps_k[t,a])"), "bound": "hbm",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback",
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": avg_ms,
                "traffic": ncu_traffic if world == 1 and name == DEFAULT_WORKLOAD else None}
    eps_bytes = 4.0 * k_local * T * A
    per_kernel = {}
    for kname, ms in kernels.items():
        d = {"ms": ms}
        if kname in ("sample", "rollout"):
            d["hbm_gbs"] = eps_bytes / (ms * 1e-3) / 1e9
            d["hbm_frac"] = d["hbm_gbs"] / peak
        per_kernel[kname] = d

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_desc(name, K, T, A), "k_local": k_local,
                   "l2": "inputs larger than L2 (eps %.0f MB per GPU per pass vs 126 MB L2)"
                         % (eps_bytes / 1e6) if eps_bytes > 200e6 else
                         "working set fits L2 (%.0f MB eps): latency-bound config, no flush" % (eps_bytes / 1e6),
                   "flags": flags, "graph": not (flags & capi.FLAG_NO_GRAPH),
                   "chain": ("one kernel: sample+rollout warps || weights+average warps -> merge+finalize"
                             if one_kernel else
                             "sample+rollout(fused) -> weights -> average -> finalize"
                             if flags & capi.FLAG_FUSED_SAMPLING else
                             "rollout -> weights -> average -> finalize; the sampler of step n+1 runs "
                             "behind step n's chain on a second stream (pipelined sampling)"
                             if flags & capi.FLAG_PIPELINED_SAMPLING else
                             "sample -> rollout -> weights -> average -> finalize"),
                   "timing": "value: CUDA events on the controller stream around K graph launches; "
                             "kernels/roofline: CUDA events between kernels in a second region of K steps"},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": 4 * 2 * A,
                "d2h_bytes_per_step": 4 * A, "ms_per_step": e2e_sec * 1e3},
        "latency_ms": {"p50": lat_ms[len(lat_ms) // 2], "p99": lat_ms[min(len(lat_ms) - 1, int(0.99 * len(lat_ms)))],
                       "min": lat_ms[0], "max": lat_ms[-1]},
        "gpu_launches": launches,
        "roofline": roofline,
        "kernels": per_kernel,
        "clocks": clk,
    }
    if one_kernel:
        per_kernel = {"step": {"ms": avg_ms, "hbm_gbs": achieved, "hbm_frac": achieved / peak,
                               "ms_with_event_pairs": kernels["average"]}}
        out["kernels"] = per_kernel
    if world == 1 and flags & (capi.FLAG_FUSED_SAMPLING | capi.FLAG_STEP_KERNEL):
        # the kernel chains timed beside it on the same workload: fused (sample+rollout, average)
        # and the canonical unfused one (sample, rollout, average)
        ctl.close()
        ctl = None
        base = flags & ~(capi.FLAG_FUSED_SAMPLING | capi.FLAG_STEP_KERNEL)
        others = {}
        for cname, cflags in (("fused_2_kernels", base | capi.FLAG_FUSED_SAMPLING), ("unfused_3_kernels", base)):
            if cflags == flags:
                continue
            c2 = m.PointMassModel(K, T, dt, 2 * A, A, seed=0, flags=cflags, device=local_rank)
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
            others[cname] = {
                "ms_per_step": ms4, "value": K * T / (ms4 * 1e-3),
                "kernels": {k: ({"ms": v, "hbm_gbs": eps_bytes / (v * 1e-3) / 1e9,
                                 "hbm_frac": eps_bytes / (v * 1e-3) / 1e9 / peak}
                                if k in ("sample", "rollout", "average") else {"ms": v})
                            for k, v in kt4.items()}}
        out["other_chains"] = others
    if world > 1:
        if args.comm == "p2p":
            out["collectives_ms"] = {
                "kind": "ONE exchange per step over NVLink peer mailboxes (direct P2P stores + flags): "
                        "every shard averages relative to its own minimum, the exchange kernel rescales "
                        "by exp(-(beta_r-beta)/lambda), sums in rank order and applies the U update",
                "min_u64": None,
                "merge_i64+finalize": kernels.get("comm_sum")}
        else:
            out["collectives_ms"] = {
                "kind": "ncclAllReduce(min) + ncclAllReduce(sum) inside the CUDA graph",
                "min_u64": kernels.get("comm_min"), "sum_i64": kernels.get("comm_sum")}
        out["config"]["comm"] = args.comm

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
                    help="MPPI_FLAG_* bits; default MPPI_FLAG_AUTO_CHAIN: >= 4e5 samples/GPU the one-kernel "
                         "step (128), >= 1.2e5 fused sampling (32), else the unfused chain with pipelined "
                         "sampling (512)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
