"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): scripts/dist_check.py under
torchrun -- row-partitioned model (fused peer-push exchange and NCCL all-gather) vs the
single-GPU model on the same graph, inputs and parameters."""
import os
import subprocess
import sys

import pytest
import torch

from helpers import ROOT

pytestmark = pytest.mark.gpu


def _run(env_extra):
    n = min(torch.cuda.device_count(), 2)
    env = dict(os.environ, **env_extra)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "scripts", "dist_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "DIST_CHECK PASS" in out, out[-3000:]
    return out


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_row_partition_fused_push_matches_single_gpu():
    out = _run({"ACMB200_PUSH": "1"})
    assert "push=True" in out or "symmetric memory unavailable" in out


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_row_partition_nccl_allgather_matches_single_gpu():
    _run({"ACMB200_PUSH": "0"})
