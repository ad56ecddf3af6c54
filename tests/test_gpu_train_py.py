"""End-to-end acceptance: the reference's OWN ACM-Pytorch/train.py, unmodified, on top of the
drop-in layer (python -m acm_gnn_b200.run).  Needs the staged reference files under
baseline/_ref (scripts/stage_reference.py; git-ignored, shipped to the GPU box) -- skipped
when they are absent."""
import os
import re
import subprocess
import sys

import pytest

from helpers import ROOT

pytestmark = pytest.mark.gpu
TRAIN = os.path.join(ROOT, "baseline", "_ref", "ACM-Pytorch", "train.py")


def _run(args, dtype):
    env = dict(os.environ, ACMB200_DTYPE=dtype, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "acm_gnn_b200.run", TRAIN] + args, cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=900)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    m = re.search(r"Test Mean: ([0-9.]+)", out)
    assert m, out[-3000:]
    return float(m.group(1)), out


@pytest.mark.skipif(not os.path.exists(TRAIN), reason="reference files not staged under baseline/_ref")
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_reference_train_py_cora_unchanged(dtype):
    """BASELINE config 1 (Cora, acmgcn, hidden 64).  The reference on CPU reaches test acc
    0.827 after 20 epochs on split 0; the paper band over 10 random
    splits is 88.62 +- 1.22 at convergence."""
    acc, out = _run(["--dataset_name", "cora", "--model", "acmgcn", "--epochs", "60", "--num_splits", "1",
                     "--fixed_splits", "1", "--hidden", "64"], dtype)
    print(f"reference train.py on the drop-in, cora acmgcn [{dtype}]: test acc {acc:.4f}")
    assert acc >= 0.80, out[-2000:]


@pytest.mark.skipif(not os.path.exists(TRAIN), reason="reference files not staged under baseline/_ref")
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_reference_train_py_squirrel_acmgcnp_structure(dtype):
    """BASELINE config 2 (Squirrel, ACM-GCN+ with structure info; published recipe
    experiment/acmgcnp_reproduce_fixed_splits.sh:6) runs unchanged through attention4.
    Acceptance band = the unmodified reference's own result for this exact command on CPU
    (torch 2.11, this container): test acc 0.4938 / 0.4938 / 0.4918 / 0.4899 for --seed 42 / 1 / 2 / 3
    after the same 40 epochs on split 0.  The dropout masks (p = 0.6) come from the CUDA generator
    here and from the CPU generator there, so the runs are not bit-comparable (measured on B200:
    0.4755 in fp32, 0.4851 in bf16): require the reference's accuracy minus 0.05."""
    acc, out = _run(["--dataset_name", "squirrel", "--model", "acmgcnp", "--structure_info", "1", "--variant", "0",
                     "--lr", "0.002", "--weight_decay", "1e-4", "--dropout", "0.6", "--epochs", "40",
                     "--num_splits", "1", "--fixed_splits", "1"], dtype)
    print(f"reference train.py on the drop-in, squirrel acmgcnp + structure [{dtype}]: test acc {acc:.4f} (reference on CPU: 0.4938)")
    assert acc >= 0.44, out[-2000:]
