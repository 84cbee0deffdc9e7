"""Fused multi-GPU exchange (dojo_step_gather_async, include/dojo_b200.h): needs >= 2 GPUs on the box (skipped otherwise).  Launches
tools/gather_check.py under torchrun with 2 ranks: the gathered buffer the step kernel fills through peer writes must equal an NCCL
all-gather of the plain per-rank results bit for bit, on every rank, for the forward and the gradient variant."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fused_gather_matches_nccl_all_gather():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29731",
           os.path.join(ROOT, "tools", "gather_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("gathered == all_gather: True") == 2, r.stdout[-2000:]
