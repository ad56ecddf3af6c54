"""The reference's own UNMODIFIED model files (ACM-Pytorch/models/models.py and
ACM-Geometric/models.py, staged under baseline/_ref) over the drop-in layer reproduce the golden
runs of the unmodified reference NUMERICALLY -- output, loss, attention, every gradient, and the
``torch.no_grad`` eval path -- not only by accuracy.  See tests/ref_model_check.py."""
import os
import subprocess
import sys

import pytest

from helpers import ROOT

pytestmark = pytest.mark.gpu
REF = os.path.join(ROOT, "baseline", "_ref")


def _run(flavour):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_model_check.py"), flavour], cwd=ROOT,
                       env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True, timeout=900)
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "REF_MODEL_CHECK PASS" in out, out[-4000:]
    return out


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "ACM-Pytorch", "models", "models.py")),
                    reason="reference files not staged under baseline/_ref")
def test_reference_pytorch_models_py_over_drop_in_matches_golden():
    out = _run("pytorch")
    assert out.count("ref_model_check[pytorch]") >= 5


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "ACM-Geometric", "models.py")),
                    reason="reference files not staged under baseline/_ref")
def test_reference_geometric_models_py_over_drop_in_matches_golden():
    """run.install("geometric"): the Geometric flavour (LayerNorm live, variant 1 default) end to end."""
    out = _run("geometric")
    assert out.count("ref_model_check[geometric]") >= 4
