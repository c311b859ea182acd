"""Sharded (one process per GPU, NCCL) run == single-GPU run of the same seed.  Needs >= 2
GPUs on the box; skipped otherwise (the driver's 1-GPU test box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_run_reproduces_single_gpu_run():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    g = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(g),
           "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(ROOT, "tools", "multigpu_check.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and "MULTIGPU_OK" in r.stdout, r.stdout[-4000:]
