"""N>1 on real GPUs (skipped unless >= 2 CUDA devices): sharded odometry == single-GPU odometry."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("comm", ["p2p", "nccl"])
@pytest.mark.parametrize("mode", ["kdtree", "projective"])
def test_sharded_matches_single_gpu(mode, comm):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "mgpu_check.py"), mode, comm]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
