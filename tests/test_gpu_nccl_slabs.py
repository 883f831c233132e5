"""The real multi-GPU path: one process per GPU, ghost rows over NCCL send/recv (NVLink). Needs >= 2
GPUs (`gpurun --gpus 2`); skipped on a single-GPU box, where tests/test_gpu_slabs.py covers the same
slab logic with the in-process transport."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("transport", ["peer", "nccl"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_slabs_equal_whole_domain(world, transport):
    """transport = peer: strips stored into the neighbour's memory by the pack kernel (CUDA IPC over NVLink,
    the default); nccl: grouped ncclSend/ncclRecv (the fallback, forced with PEDONI_SLAB_TRANSPORT=nccl)."""
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    import os
    env = dict(os.environ)
    if transport == "nccl":
        env["PEDONI_SLAB_TRANSPORT"] = "nccl"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world + (10 if transport == "nccl" else 0)),
           str(ROOT / "tests" / "nccl_slab_worker.py"), "400000", "25"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0 and "NCCL-SLABS OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    assert ("peer-memory" if transport == "peer" else "nccl send/recv") in r.stdout, r.stdout[-500:]


def test_multi_gpu_slabs_on_a_growing_scenario():
    """bottleneck.toml for 400 ticks on 2 GPUs (peer-memory transport, device-side spawning, buffers growing,
    nearly everybody in the slab that holds the gap) = whole domain, bit for bit."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29640",
           str(ROOT / "tests" / "nccl_slab_worker.py"), "scenario:bottleneck", "400"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "NCCL-SLABS OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
