"""Needs >= 2 GPUs on the box (skips otherwise): slab decomposition + NCCL halo exchange parity."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_slabs_match_oracle():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tests", "multigpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "MULTIGPU_CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
