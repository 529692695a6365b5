"""Multi-GPU parity under pytest: launches tests/nccl_check.py with torchrun on every GPU of the box (2..8) over NCCL.
Self-skips on a box with fewer than two devices (the single-GPU driver box); `gpurun --gpus N` runs it for real."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_nccl_sharded_run_is_identical_to_one_gpu():
    """broadcast set + sharded clean (both modes), sharded diff, and the one-pass device / host planes: the rank-order
    concatenation equals the single-GPU output byte for byte, counters summed (tests/nccl_check.py)"""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = min(n, 8)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                        os.path.join(ROOT, "tests", "nccl_check.py")], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert "nccl_check ok" in r.stdout and "MISMATCH" not in r.stdout
