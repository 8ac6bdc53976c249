"""Data-parallel step on REAL GPUs (needs >= 2 visible devices; the single-GPU driver run skips it, the builder's
`gpurun --gpus 2` run is recorded under profiles/): launches tests/dist_check.py under torchrun."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_allreduced_gradients_and_replica_consistency(lib):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "dist_check.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "DIST_CHECK_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


def test_peer_exchange_two_processes_one_gpu(lib):
    """The peer-memory gradient exchange (csrc/peer.cu) between two PROCESSES sharing cuda:0 (CUDA IPC works across
    processes on one device; the GPU time-slices the two contexts, so the flag waits cross a context switch): bit-exact
    rank-ordered mean, replicas identical after six captured steps on different batches."""
    env = dict(os.environ, KP_DIST_SAME_GPU="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29612", os.path.join(ROOT, "tests", "dist_check.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=400, env=env)
    assert out.returncode == 0 and "DIST_CHECK_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
    assert "exchange=peer-memory kernel" in out.stdout, out.stdout[-2000:]
