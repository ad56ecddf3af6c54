"""Host-side logic of the drop-in module on CPU: interface parity with the reference class
(names, shapes, order, RNG draw order), parameter packing, branch selection, and the
absence of any CPU fallback."""
import numpy as np
import pytest
import torch

from helpers import Golden

import acm_gnn_b200 as A
from acm_gnn_b200 import functional as Fn
from acm_gnn_b200 import layers_geometric


def test_padded_width():
    assert [Fn.padded_width(f) for f in (1, 2, 5, 7, 8, 9, 16, 24, 40, 64, 65, 128, 200, 256)] == \
        [8, 8, 8, 8, 8, 16, 16, 32, 64, 64, 128, 128, 256, 256]
    with pytest.raises(NotImplementedError):
        Fn.padded_width(257)


def test_pack_layout_matches_header():
    fp, f = 16, 5
    a = [torch.arange(f, dtype=torch.float32).reshape(f, 1) + 10 * k for k in range(4)]
    av = torch.arange(16, dtype=torch.float32).reshape(4, 4)
    ln = [(torch.full((f,), 2.0 + k), torch.full((f,), -1.0 - k)) for k in range(4)]
    p = Fn.build_pack(fp, f, a, av, ln)
    assert p.numel() == 16 * fp + 32  # ACM_PACK_FLOATS_VALUE: 12fp+16 parameters + derived LayerNorm tail
    for k in range(4):
        assert torch.equal(p[k * fp:k * fp + f], a[k].reshape(-1))
        assert float(p[k * fp + f:(k + 1) * fp].abs().sum()) == 0.0  # zero padding
        assert torch.equal(p[4 * fp + 16 + k * fp:4 * fp + 16 + k * fp + f], ln[k][0])
        assert torch.equal(p[8 * fp + 16 + k * fp:8 * fp + 16 + k * fp + f], ln[k][1])
    assert torch.equal(p[4 * fp:4 * fp + 16].view(4, 4), av)
    for k in range(4):  # derived tail: gamma*a and the per-channel sums
        assert torch.allclose(p[12 * fp + 16 + k * fp:12 * fp + 16 + k * fp + f], ln[k][0] * a[k].reshape(-1))
        assert float(p[16 * fp + 16 + k]) == pytest.approx(float((ln[k][1] * a[k].reshape(-1)).sum()))
        assert float(p[16 * fp + 20 + k]) == pytest.approx(float((ln[k][0] * a[k].reshape(-1)).sum()))
    # 3x3 att_vec lands in the top-left of the 4x4 slot block
    p3 = Fn.build_pack(fp, f, a[:3], av[:3, :3].contiguous(), None)
    assert torch.equal(p3[4 * fp:4 * fp + 16].view(4, 4)[:3, :3], av[:3, :3])
    assert float(p3[4 * fp:4 * fp + 16].view(4, 4)[3].abs().sum()) == 0.0


def test_wcat_layout():
    fp, f, fin = 8, 3, 4
    ws = [torch.full((fin, f), float(k + 1)) for k in range(3)]
    w = Fn.build_wcat(fp, f, ws, torch.float32)
    assert w.shape == (fin, 3 * fp)
    for k in range(3):
        assert float(w[:, k * fp:k * fp + f].min()) == k + 1 and float(w[:, k * fp + f:(k + 1) * fp].abs().sum()) == 0


@pytest.mark.parametrize("name", ["gcn_pt_acmgcn_v0", "gcn_pt_acmgcnpp_v1_s1", "gcn_geo_acmgcnp_v0_s1"])
def test_state_dict_contract_and_same_seed_init(name):
    """Parameter names, shapes and creation order equal the reference's state_dict; same-seed
    init is bitwise identical (RNG draw order of layers.py:70-92 and of GCN.__init__)."""
    g = Golden(name)
    torch.manual_seed(g.seed)
    _ = torch.rand(g.n, g.nfeat); _ = torch.randint(0, g.nclass, (g.n,)); _ = torch.randperm(g.n)
    model = A.GCN(g.nfeat, g.nhid, g.nclass, 2, g.n, 0.0, g.model_type, g.structure_info,
                  variant=bool(g.variant), flavour=g.flavour)
    sd = model.state_dict()
    ref_keys = [k[len("param/"):] for k in g.z.files if k.startswith("param/")]
    ours = [k for k in sd if k not in ("fea_param", "xX_param") and ".bns." not in k]
    assert sorted(ours) == sorted(ref_keys)
    for k in ref_keys:
        assert tuple(sd[k].shape) == g.z["param/" + k].shape, k
        assert np.array_equal(sd[k].numpy(), g.z["param/" + k]), f"same-seed init differs for {k}"
    layer = model.gcns[0]
    assert repr(layer) == f"GraphConvolution ({g.nfeat} -> {g.nhid})"
    for attr in ("in_features", "out_features", "output_layer", "model_type", "structure_info", "variant"):
        assert hasattr(layer, attr)


def test_parameter_order_matches_reference_class():
    ref_order = ["weight_low", "weight_high", "weight_mlp", "att_vec_low", "att_vec_high", "att_vec_mlp",
                 "layer_norm_low.weight", "layer_norm_low.bias", "layer_norm_high.weight", "layer_norm_high.bias",
                 "layer_norm_mlp.weight", "layer_norm_mlp.bias", "layer_norm_struc_low.weight",
                 "layer_norm_struc_low.bias", "layer_norm_struc_high.weight", "layer_norm_struc_high.bias",
                 "att_struc_low", "struc_low", "att_vec"]
    layer = A.GraphConvolution(6, 4, 10, "acmgcn")
    # nn.Module lists direct parameters first, then sub-module parameters -- as the reference does
    direct = [n for n, _ in layer.named_parameters() if "." not in n]
    assert direct == [n for n in ref_order if "." not in n]
    assert layer.att_vec.shape == (3, 3)
    assert A.GraphConvolution(6, 4, 10, "acmgcnp", structure_info=1).att_vec.shape == (4, 4)


def test_layernorm_quirk_q1_and_structure_rules():
    pt = A.GraphConvolution(6, 4, 10, "acmgcnp")
    geo = layers_geometric.GraphConvolution(6, 4, 10, "acmgcnp")
    assert pt._ln_live() is False and geo._ln_live() is True
    assert A.GraphConvolution(6, 4, 10, "acmgcn+")._ln_live() is True
    assert layers_geometric.GraphConvolution(6, 4, 10, "acmgcn")._ln_live() is False
    assert A.GraphConvolution(6, 4, 10, "acmgcnp", structure_info=1)._uses_structure()
    assert not A.GraphConvolution(6, 4, 10, "acmgcn", structure_info=0)._uses_structure()


def test_lazy_struc_low_only_when_unused_and_huge(monkeypatch):
    from acm_gnn_b200 import layers as L
    monkeypatch.setattr(L, "_LAZY_STRUC_ELEMS", 100)
    assert L.GraphConvolution(3, 4, 1000, "acmgcn").struc_low.shape == (0, 4)
    assert L.GraphConvolution(3, 4, 1000, "acmgcnp", structure_info=1).struc_low.shape == (1000, 4)
    assert L.GraphConvolution(3, 4, 10, "acmgcn").struc_low.shape == (10, 4)


def test_no_cpu_fallback():
    """CPU tensors are rejected loudly: the product path has no CPU / eager fallback."""
    layer = A.GraphConvolution(6, 4, 10, "acmgcn")
    x = torch.rand(10, 6)
    adj = torch.eye(10)
    with pytest.raises(RuntimeError, match="CUDA"):
        layer(x, adj, adj.to_sparse(), None)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from acm_gnn_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.AcmLibraryError, match="no CPU fallback"):
        _lib.load()


def test_product_does_not_import_oracle():
    import os, re
    root = os.path.dirname(os.path.abspath(A.__file__))
    for fn in os.listdir(root):
        if fn.endswith(".py"):
            src = open(os.path.join(root, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn


def test_graphed_step_has_no_cpu_path():
    """The CUDA-graph wrappers refuse CPU tensors like every other entry point (no fallback)."""
    from acm_gnn_b200.graphed import GraphedForward, GraphedTrainStep
    model = A.GCN(6, 8, 3, 2, 10, 0.0, "acmgcn", 0)
    x = torch.rand(10, 6)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        GraphedForward(model, x, (None, None, None))
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        GraphedTrainStep(model, opt, x, (None, None, None), torch.zeros(10, dtype=torch.int64), torch.ones(10, dtype=torch.uint8))


def test_row_order_groups_equal_degrees_inside_windows(monkeypatch):
    """CsrMatrix.row_order: a permutation, degree-sorted (stable) inside each window of consecutive rows."""
    from acm_gnn_b200.operator import CsrMatrix
    monkeypatch.setenv("ACMB200_ROW_ORDER", "1")
    g = torch.Generator().manual_seed(0)
    n = 10000
    deg = torch.randint(0, 40, (n,), generator=g)
    rowptr = torch.zeros(n + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(deg, 0)
    m = CsrMatrix(n, n, rowptr, torch.zeros(int(rowptr[-1]), dtype=torch.int32), torch.zeros(int(rowptr[-1])))
    order = m.row_order().long()
    assert torch.equal(torch.sort(order).values, torch.arange(n))
    w = m.ORDER_WINDOW
    for s in range(0, n, w):
        blk = order[s:s + w]
        assert int(blk.min()) >= s and int(blk.max()) < min(n, s + w)
        d = deg[blk]
        assert bool((d[1:] >= d[:-1]).all())
        same = d[1:] == d[:-1]
        assert bool((blk[1:][same] > blk[:-1][same]).all())      # stable: ties keep the natural order
    assert m.row_order() is m.row_order()                        # cached
