"""Runs tests/dist_gpu_check.py under torchrun when at least two GPUs are visible (NCCL, one process per GPU)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_paths_match_single_gpu_over_nccl():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (the gloo world_size-2 tests cover the host logic on CPU)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(8, n)),
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(REPO, "tests", "dist_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=REPO)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    rep = json.loads(line)
    assert rep["ok"] is True and rep["world"] == min(8, n)
