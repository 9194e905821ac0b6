"""x-slab decomposition across GPUs (NCCL halo exchange + distributed FFT transposes) vs the single-GPU run."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("p2p", [1, 0])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_slabs_match_single_gpu(world, p2p):
    """p2p = 1: ghost cells and FFT transposes as CUDA-IPC peer loads; p2p = 0: NCCL send/recv."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + 10 * world + p2p), os.path.join(ROOT, "scripts", "multi_gpu_check.py"), "64", "3"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=180, env=dict(os.environ, BZ_P2P=str(p2p)))
    assert "MULTI_GPU_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("p2p", [1, 0])
def test_bomex_slabs_match_single_gpu(p2p):
    """BOMEX-type physics on 2 slabs: the subsidence forcing's horizontal means are all-reduced across ranks every stage."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29650 + p2p), os.path.join(ROOT, "scripts", "multi_gpu_check.py"), "64", "3", "bomex"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=180, env=dict(os.environ, BZ_P2P=str(p2p)))
    assert "MULTI_GPU_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
