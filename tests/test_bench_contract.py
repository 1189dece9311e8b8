"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the
required keys, and the product arm refuses to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def test_reference_arm_prints_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--steps", "1", "--warmup", "0", "--workload", "point_mass2d_K1e4_T200"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "rollout_steps_per_s"
    assert d["unit"] == "rollout-steps/s" and d["higher_is_better"] is True
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data",
                "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["value"] > 0 and "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stdout + r.stderr)


@pytest.mark.gpu
def test_product_arm_prints_contract_line():
    """One short run of the product arm on the GPU: the JSON line carries every key of the
    bench contract (value, e2e with host<->device bytes, roofline, cpu_baseline, clocks,
    gpu_launches) and names the workload."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "5", "--warmup", "3",
                        "--workload", "point_mass2d_K1e5_T200"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["metric"] == "rollout_steps_per_s" and d["unit"] == "rollout-steps/s"
    assert d["n_gpus"] == 1 and d["steps"] == 5 and d["higher_is_better"] is True
    assert d["value"] > 0 and d["gpu_launches"] >= 5 and "workload" in d["config"]
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 16 and d["e2e"]["d2h_bytes_per_step"] == 8
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 1 and cb["value"] > 0
    assert "sm_mhz" in d["clocks"] and "reasons" in d["clocks"]
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and d["scaling"] == "strong"


def test_quoted_dram_traffic_is_not_stale():
    """bench.py quotes `roofline.traffic` from profiles/*_traffic.json (ncu captures): every file
    names the commit it was captured at and the kernel sources it depends on, and none of those
    sources may have changed since (re-capture with tools/jobs/r2b_final1.sh).  Needs the git
    history: skipped where the tree travels without it (the GPU boxes)."""
    import json
    import subprocess
    from conftest import ROOT
    if not os.path.isdir(os.path.join(ROOT, ".git")):
        pytest.skip("no git history here")

    def git(*a):
        return subprocess.run(["git", "-C", ROOT] + list(a), capture_output=True, text=True)

    for name in ("step_traffic.json", "average_traffic.json", "tile_traffic.json"):
        d = json.load(open(os.path.join(ROOT, "profiles", name)))
        assert d.get("commit") and d.get("kernel_sources"), name
        if git("cat-file", "-e", d["commit"] + "^{commit}").returncode != 0:
            pytest.skip(f"commit {d['commit']} of profiles/{name} is not in this history (shallow clone?)")
        for src in d["kernel_sources"]:
            last = git("log", "-1", "--format=%H", "--", src).stdout.strip()
            # the capture's commit must contain the last change of the source
            assert git("merge-base", "--is-ancestor", last, d["commit"]).returncode == 0, \
                f"{src} changed after the capture quoted in profiles/{name}"
