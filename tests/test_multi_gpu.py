"""GPU: the NCCL-sharded large BA (one process per GPU; world size 1 on a one-GPU box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_ba_ranks():
    """Two ranks where the box has two GPUs; on a one-GPU box the same script runs the sharded code
    path (NCCL communicator, separate control kernel, collective loop exit) at world size 1."""
    import torch
    n = min(2, torch.cuda.device_count())
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", "29613",
           os.path.join(ROOT, "tests", "multi_gpu", "sharded_ba.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SHARDED_BA_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
