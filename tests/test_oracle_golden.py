"""The oracle (oracle/acm_oracle.py) against the golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, Golden, O, flat_param_name, golden_cases, oracle_run

ATOL, RTOL = 2e-6, 2e-5  # fp32 restatement vs fp32 reference; only summation order differs


@pytest.mark.parametrize("name", golden_cases())
def test_operator_bit_exact(name):
    """CSR indices and degree-normalised weights are BIT-EXACT with the reference's
    operators (ACM-Pytorch/utils.py:421-438,626-628; ACM-Geometric/utils.py:5-28)."""
    g = Golden(name)
    op = g.operator()
    z = g.z
    assert np.array_equal(op.rowptr, z["low_crow"])
    assert np.array_equal(op.col, z["low_col"])
    assert np.array_equal(op.w_low.view(np.uint32), z["low_val"].view(np.uint32))
    keep = op.w_high != 0
    hi_idx = np.stack([op.rows()[keep], op.col[keep]])
    assert np.array_equal(hi_idx, z["high_idx"])
    assert np.array_equal(op.w_high[keep].view(np.uint32), z["high_val"].view(np.uint32))


@pytest.mark.parametrize("ds", ["cora", "squirrel"])
def test_dataset_operator_bit_exact(ds):
    """Same on the two fixture graphs of BASELINE configs 1-2 (Squirrel carries 140 data
    self-loops -> diagonal multiplicity 2, quirk Q5)."""
    z = np.load(os.path.join(GOLDEN, f"dataset_{ds}.npz"))
    n = int(z["n"])
    op = O.build_operator(z["row"].astype(np.int64), z["col"].astype(np.int64), n, "pytorch", val=z["raw_val"])
    assert np.array_equal(op.rowptr, z["low_crow"].astype(np.int64))
    assert np.array_equal(op.col, z["low_col"].astype(np.int64))
    assert np.array_equal(op.w_low.view(np.uint32), z["low_val"].view(np.uint32))
    keep = op.w_high != 0
    assert np.array_equal(np.stack([op.rows()[keep], op.col[keep]]), z["high_idx"].astype(np.int64))
    assert np.array_equal(op.w_high[keep].view(np.uint32), z["high_val"].view(np.uint32))
    if ds == "squirrel":
        assert op.nnz == 401907 and int((op.mult == 2).sum()) == 140
    else:
        assert op.nnz == 13264


@pytest.mark.parametrize("name", golden_cases())
def test_forward_backward_matches_reference(name):
    g = Golden(name)
    out, atts, loss, gx, p = oracle_run(g)
    z = g.z
    np.testing.assert_allclose(out.detach().numpy(), z["out"], atol=ATOL, rtol=RTOL)
    np.testing.assert_allclose(float(loss), float(z["loss"]), atol=ATOL, rtol=RTOL)
    for li, a in enumerate(atts):
        np.testing.assert_allclose(a.detach().numpy(), z[f"att{li}"], atol=ATOL, rtol=RTOL)
    np.testing.assert_allclose(gx.numpy(), z["grad_x"], atol=ATOL, rtol=1e-4)
    ref_grads = g.grads()
    checked = 0
    for grp, d in p.items():
        for sub, t in d.items():
            rg = ref_grads[flat_param_name(grp, sub)]
            if rg.size == 0:  # parameter unused on this path in the reference (no grad)
                assert t.grad is None or float(t.grad.abs().max()) == 0.0, (grp, sub)
                continue
            assert t.grad is not None, (grp, sub)
            np.testing.assert_allclose(t.grad.numpy(), rg, atol=ATOL, rtol=1e-4, err_msg=f"{grp}.{sub}")
            checked += 1
    assert checked >= 7


def test_init_draw_order_matches_reference():
    """Same-seed parameter init reproduces the stored reference state_dict bitwise
    (draw order of ACM-Pytorch/models/layers.py:70-92)."""
    g = Golden("gcn_pt_acmgcn_v0")
    # make_golden seeds, then draws: graph (numpy rng), x, labels, randperm, then GCN init
    torch.manual_seed(g.seed)
    _ = torch.rand(g.n, g.nfeat)
    _ = torch.randint(0, g.nclass, (g.n,))
    _ = torch.randperm(g.n)
    p = O.init_gcn_params(g.nfeat, g.nhid, g.nclass, g.n, g.model_type, g.structure_info)
    for grp in ("gcns.0", "gcns.1"):
        for sub in ("weight_low", "weight_high", "weight_mlp", "att_vec_low", "att_vec_high", "att_vec_mlp", "att_vec"):
            assert np.array_equal(p[grp][sub].numpy(), g.z[f"param/{grp}.{sub}"]), (grp, sub)
