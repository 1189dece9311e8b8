"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports
every symbol include/mppi_b200.h declares, and fails loudly (no fallback) when no
device is present."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from mppi_gpu_b200 import capi


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "mppi_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mppi_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    names = _declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"libmppi_b200.so does not export {n}"
    assert sorted(capi.EXPORTS) == names, "capi.EXPORTS out of sync with the header"


def test_abi_version_and_struct_layout():
    lib = capi.load()
    assert lib.mppi_abi_version() == 3
    p = capi.MppiParams()
    assert lib.mppi_params_default(C.byref(p)) == 0
    # the C side wrote sizeof(mppi_params): the ctypes mirror must agree
    assert p.struct_size == C.sizeof(capi.MppiParams)
    # reference-compatible preset (src/point_mass.cu:53-54, src/point_mass_gpu.cu:58-61,86)
    assert p.lambda_ == 1.0 and p.world_size == 1 and p.comm == capi.COMM_NONE
    assert all(abs(p.sigma[a] - 0.025) < 1e-9 for a in range(capi.MAX_ACT))
    assert all(p.inv_sigma[a] == 1.0 for a in range(capi.MAX_ACT))
    assert p.flags == 0 and p.seed == 0 and p.model == capi.MODEL_POINT_MASS


def test_kernel_names():
    lib = capi.load()
    names = [lib.mppi_kernel_name(i).decode() for i in range(capi.K_COUNT)]
    assert names == ["sample", "rollout", "comm_min", "weights", "average", "comm_sum", "finalize"]


def test_shard_ranges_tile_the_samples():
    for K in (1, 3, 4, 5, 1000, 10007, 1000000, 1000003):
        for W in (1, 2, 3, 4, 8):
            if (K + 3) // 4 < W:
                continue
            edges = [capi.shard_range(K, r, W) for r in range(W)]
            assert edges[0][0] == 0 and edges[-1][1] == K
            for (b0, e0), (b1, e1) in zip(edges, edges[1:]):
                assert e0 == b1 and b1 % 4 == 0          # whole Philox quads
            sizes = [e - b for b, e in edges]
            assert max(sizes) - min(sizes) <= 4 + 3
    with pytest.raises(capi.MppiError):
        capi.shard_range(10, 3, 2)


def test_invalid_parameters_are_rejected():
    lib = capi.load()
    p = capi.MppiParams()
    lib.mppi_params_default(C.byref(p))
    h = C.c_void_p()
    p.samples, p.horizon, p.dt, p.act_dim, p.state_dim = 100, 10, 0.1, 2, 5
    assert lib.mppi_create(C.byref(p), C.byref(h)) == capi.ERR_INVALID
    assert b"state_dim" in lib.mppi_last_error()
    p.state_dim, p.act_dim = 10, 5
    assert lib.mppi_create(C.byref(p), C.byref(h)) == capi.ERR_INVALID
    p.state_dim, p.act_dim, p.struct_size = 4, 2, 12
    assert lib.mppi_create(C.byref(p), C.byref(h)) == capi.ERR_INVALID
    assert b"ABI" in lib.mppi_last_error()


def test_no_silent_cpu_fallback():
    """Without a CUDA device mppi_create must fail with MPPI_ERR_NO_DEVICE."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = capi.load()
    p = capi.MppiParams()
    lib.mppi_params_default(C.byref(p))
    p.samples, p.horizon, p.dt, p.act_dim, p.state_dim = 100, 10, 0.1, 2, 4
    h = C.c_void_p()
    assert lib.mppi_create(C.byref(p), C.byref(h)) == capi.ERR_NO_DEVICE
    assert not h.value
    assert b"no CPU fallback" in lib.mppi_last_error()


def test_product_does_not_touch_the_oracle():
    """Nothing under mppi_gpu_b200/, include/ or cpp/ may reference oracle/."""
    bad = []
    for top in ("mppi_gpu_b200", "include", "cpp"):
        for dp, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
                    txt = open(os.path.join(dp, f), errors="replace").read()
                    if re.search(r"(import|from)\s+oracle|oracle/|liboracle|pyoracle", txt):
                        # comments that merely name the oracle's restatement are fine
                        hits = [l for l in txt.splitlines()
                                if re.search(r"(import|from)\s+oracle|liboracle|pyoracle|#include.*oracle", l)]
                        if hits:
                            bad.append((os.path.join(dp, f), hits))
    assert not bad, bad


def test_header_is_plain_c_and_links(tmp_path):
    """include/mppi_b200.h is what a cgo / FFI / C caller binds: it must compile as strict C99
    and a C program must link against the library without any C++ or CUDA header."""
    import subprocess
    src = tmp_path / "bind.c"
    src.write_text('#include "mppi_b200.h"\n'
                   "int main(void) { mppi_params p; if (mppi_params_default(&p)) return 2;\n"
                   "  return (p.struct_size == sizeof p && mppi_abi_version() == MPPI_ABI_VERSION) ? 0 : 1; }\n")
    exe = tmp_path / "bind"
    libdir = os.path.join(ROOT, "mppi_gpu_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", str(src),
                    "-I", os.path.join(ROOT, "include"), "-L", libdir, "-lmppi_b200",
                    f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_auto_chain_policy_against_the_recorded_sweep():
    """MPPI_FLAG_AUTO_CHAIN decides from a cost model (controller.cu: chain_cost), exposed as
    mppi_chain_estimate (host arithmetic).  Against the recorded sweep of the three chains on one
    B200 (profiles/r02_chain_sweep_v2.jsonl: 7 (A, T) shapes x 16 shard sizes): the chosen chain's
    measured time is within 5 % (+1.5 us) of the fastest on every shape, and the estimate of the
    chosen chain is within 30 % of its measurement from 1e5 samples up."""
    import json
    path = os.path.join(ROOT, "profiles", "r02_chain_sweep_v2.jsonl")
    rows = [json.loads(l) for l in open(path)]
    assert len(rows) >= 100
    fam = {capi.FLAG_PIPELINED_SAMPLING: "unfused", capi.FLAG_FUSED_SAMPLING: "fused",
           capi.FLAG_STEP_KERNEL: "step"}
    for r in rows:
        est, choice = capi.chain_estimate(r["K"], r["T"], r["A"], 148)
        ms = {k: v for k, v in r["ms"].items() if k in ("unfused", "fused", "step")}
        best = min(ms.values())
        got = ms[fam[choice]]
        assert got <= 1.05 * best + 1.5e-3, (r["A"], r["T"], r["K"], fam[choice], ms)
        if r["K"] >= 100000:
            k = fam[choice]
            assert abs(est[k] / (got * 1e3) - 1.0) < 0.30, (r["A"], r["T"], r["K"], k, est[k], got)
    # clear cases (the GPU test test_auto_chain_picks_by_work asks the handle for the same)
    assert capi.chain_estimate(30000, 200, 2)[1] == capi.FLAG_PIPELINED_SAMPLING
    assert capi.chain_estimate(150000, 200, 2)[1] == capi.FLAG_FUSED_SAMPLING
    assert capi.chain_estimate(450000, 10, 2)[1] == capi.FLAG_FUSED_SAMPLING
    assert capi.chain_estimate(700000, 120, 2)[1] == capi.FLAG_STEP_KERNEL
    # the rollout cost is a step function of the warps on the fullest sub-partition
    e1, _ = capi.chain_estimate(150000, 200, 3)
    e2, _ = capi.chain_estimate(166667, 200, 3)
    assert e2["fused"] - e1["fused"] > 40.0
