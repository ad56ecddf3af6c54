"""Size-independent properties of the path, checked on the CPU oracle (the same identities the
CUDA kernels rely on, and the ones bench.py re-checks at full size on the GPU):
  * A_low = D^-1 (A + I) is row-stochastic and A_high = I - A_low (ACM-Pytorch/utils.py:626-628);
  * aggregate-first order: A (X W) = (A X) W and X W_H - A (X W_H) = (X - A X) W_H;
  * variant 0: every channel output is a relu and the attention weights are a softmax, so the
    layer output is >= 0 and the relu between the layers (models.py:160) is the identity;
  * the attention columns of every layer sum to 1 per node (layers.py:118)."""
import numpy as np
import pytest
import torch

from helpers import O


def _graph(n=400, e=3000, seed=3, self_loops=True):
    row, col = O.synthetic_edges(n, e, seed=seed)
    if self_loops:  # data self-loops give the diagonal multiplicity 2 (quirk Q5)
        extra = np.arange(0, n, 37, dtype=row.dtype)
        row, col = np.concatenate([row, extra]), np.concatenate([col, extra])
    return row, col, n


@pytest.mark.parametrize("flavour", ["pytorch", "geometric"])
def test_low_pass_operator_is_row_stochastic_and_high_is_its_complement(flavour):
    row, col, n = _graph()
    op = O.build_operator(row, col, n, flavour)
    low, high = O.operator_to_torch(op)
    ones = torch.ones(n, 1)
    assert torch.allclose(torch.sparse.mm(low, ones), ones, atol=2e-6)
    assert float(torch.sparse.mm(high, ones).abs().max()) < 2e-6
    dense = low.to_dense() + high.to_dense()
    assert torch.allclose(dense, torch.eye(n), atol=1e-6)
    # duplicates summed, diagonal always present
    assert bool((low.to_dense().diagonal() > 0).all())


def test_aggregate_first_order_is_the_same_map():
    row, col, n = _graph()
    op = O.build_operator(row, col, n, "pytorch")
    low, _ = O.operator_to_torch(op)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(n, 24, generator=g, dtype=torch.float64)
    wl = torch.randn(24, 16, generator=g, dtype=torch.float64)
    wh = torch.randn(24, 16, generator=g, dtype=torch.float64)
    a = low.to(torch.float64)
    z = torch.sparse.mm(a, x)
    assert torch.allclose(torch.sparse.mm(a, x @ wl), z @ wl, rtol=1e-12, atol=1e-12)
    assert torch.allclose(x @ wh - torch.sparse.mm(a, x @ wh), (x - z) @ wh, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("model_type,structure_info", [("acmgcn", 0), ("acmgcnp", 1)])
def test_variant0_output_is_nonnegative_and_attention_is_a_distribution(model_type, structure_info):
    row, col, n = _graph(self_loops=False)
    op = O.build_operator(row, col, n, "pytorch")
    low, high = O.operator_to_torch(op, dense_low=True)
    un = O.raw_adjacency_to_torch(row, col, n) if structure_info else None
    gp = torch.Generator().manual_seed(42)
    fin, hid, ncls = 20, 16, 5
    params = O.init_gcn_params(fin, hid, ncls, n if structure_info else 0, model_type, structure_info, gp)
    x = O.row_normalise_features(torch.rand(n, fin, generator=gp))
    p0 = params["gcns.0"]
    y, att = O.layer_forward(p0, x, low, high, un, model_type=model_type, variant=False,
                             structure_info=structure_info, flavour="pytorch")
    assert float(y.min()) >= 0.0                       # relu(fea1) is the identity (models.GCN.skip_identity_relu)
    assert att.shape[1] == (4 if structure_info else 3)
    assert torch.allclose(att.sum(1), torch.ones(n), atol=1e-6) and float(att.min()) > 0.0
