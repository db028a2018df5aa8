"""REAL multi-rank parity: torch.distributed.run with one process per GPU over NCCL; the sharded
engine's output on every rank must equal the single-GPU result bit for bit (tests/multirank_worker.py).  Needs at least
two GPUs in the box; the emulated-rank tests in test_parity_gpu.py cover the same arithmetic on one."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, size, layouts, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multirank_worker.py"), str(size), layouts]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-4000:] + out.stderr[-4000:]
    assert "0 failures" in out.stdout


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def test_two_ranks_equal_single_gpu(built):
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, 128, "contiguous,peer,interleaved,replica", 29631)


def test_all_ranks_equal_single_gpu(built):
    g = _gpus()
    if g < 4:
        pytest.skip("needs 4+ GPUs")
    _run(8 if g >= 8 else 4, 256, "contiguous,peer,replica", 29632)
