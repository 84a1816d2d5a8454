"""Data-parallel correctness over NCCL on real GPUs (needs >= 2 devices: `gpurun --gpus 2`; skipped on one GPU).
The worker (tests/dp_worker.py) asserts that 2 ranks x 2048 samples reproduce 1 rank x 4096 -- gradients and post-Adam
parameters, both supervision branches, graph replay on, with and without the overlapped bucket schedule."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_reproduce_one_rank():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(ROOT, "tests", "dp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    sys.stdout.write(r.stdout[-6000:])
    sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0 and "DP_CHECK PASSED" in r.stdout
