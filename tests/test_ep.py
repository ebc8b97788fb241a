"""Expert-parallel tests: world-size-2 gloo run of the host logic on CPU; NCCL run on >= 2 GPUs when present."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "ep_worker.py")


def _torchrun(nproc, mode, port, extra=(), env_extra=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), WORKER, "--mode", mode, *extra]
    env = dict(os.environ, OMP_NUM_THREADS="2", **(env_extra or {}))
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)


def test_ep_host_logic_gloo_world2():
    res = _torchrun(2, "cpu", 29611)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("cpu ep host logic ok") == 2, res.stdout[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("transport", ["nccl", "peer"])
def test_ep_matches_local_experts_nccl(transport):
    """EP layer against the same layer with every expert local, on >= 2 GPUs, through both row transports: NCCL all-to-all
    and the peer-memory kernels (every producer stores its rows into the consuming rank's buffer over NVLink, ab_ep_*)."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    res = _torchrun(world, "gpu", {"nccl": 29612, "peer": 29613}[transport], env_extra={"APERTIS_B200_EP": transport})
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("gpu ep ok") == world
