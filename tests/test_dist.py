"""Multi-rank tests: world_size-2 gloo on CPU (exchange logic), NCCL on >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

WORKER = os.path.join(ROOT, "tests", "dist_worker.py")


def _launch(mode, nproc, port):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER, mode]
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=240)


def test_k_shard_exchange_logic_gloo_world2():
    r = _launch("gloo", 2, 29541)
    assert r.returncode == 0 and "GLOO_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_k_shard_exchange_logic_gloo_world3():
    r = _launch("gloo", 3, 29542)
    assert r.returncode == 0 and "GLOO_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.gpu
@pytest.mark.parametrize("comm", ["nccl", "p2p"])
def test_k_sharded_controller(comm):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    r = _launch(comm, min(n, 4), 29543 + (comm == "p2p"))
    assert r.returncode == 0 and comm.upper() + "_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
