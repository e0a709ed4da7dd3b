"""N>1 on real GPUs (skipped unless >= 2 CUDA devices): sharded odometry == single-GPU odometry."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _check(mode, comm, rule):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "mgpu_check.py"), mode, comm, rule]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(r.stdout[-600:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("comm", ["p2p", "nccl"])
@pytest.mark.parametrize("mode", ["kdtree", "projective"])
def test_sharded_matches_single_gpu(mode, comm):
    _check(mode, comm, "fixed")


@pytest.mark.parametrize("comm", ["p2p", "nccl"])
def test_sharded_with_the_default_stop_rule(comm):
    """Exchange rounds skipped by converged frames (device-side no-op launches) must not desynchronise the ranks."""
    _check("kdtree", comm, "stop")
