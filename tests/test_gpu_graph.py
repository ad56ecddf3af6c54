"""GPU tests of the CUDA-graph captured train step / eval forward (acm_gnn_b200/graphed.py) and of
the narrow-row gather hint.  The captured step must be the eager step: same gradients after the
first replay, same loss trajectory, and the warm-up must leave no trace in the parameters."""
import os

import pytest
import torch

from helpers import Golden, O

pytestmark = pytest.mark.gpu


def _model_from_golden(g, mode):
    import acm_gnn_b200 as A
    os.environ["ACMB200_DTYPE"] = mode
    torch.manual_seed(g.seed)
    model = A.GCN(g.nfeat, g.nhid, g.nclass, 2, g.n, 0.0, g.model_type, g.structure_info,
                  variant=bool(g.variant), flavour=g.flavour).cuda()
    sd = {k[len("param/"):]: torch.from_numpy(g.z[k]) for k in g.z.files if k.startswith("param/")}
    model.load_state_dict(sd, strict=False)
    return model


def _train_params(model):
    return [p for k, p in model.named_parameters() if k not in ("fea_param", "xX_param")]


def _inputs(g):
    import acm_gnn_b200 as A
    op = A.AcmOperator.from_edges(torch.from_numpy(g.row).cuda(), torch.from_numpy(g.col).cuda(), g.n, g.flavour,
                                  with_raw=bool(g.structure_info))
    mask = torch.zeros(g.n, dtype=torch.uint8, device="cuda")
    mask[g.idx_train.cuda()] = 1
    return op, g.x.cuda(), g.labels.cuda(), mask


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["gcn_pt_acmgcn_v0", "gcn_geo_acmgcnp_v1", "gcn_pt_acmgcnpp_v1_s1"])
def test_graphed_train_step_matches_eager(name, mode):
    from acm_gnn_b200.functional import nll_log_softmax
    from acm_gnn_b200.graphed import GraphedTrainStep
    g = Golden(name)
    op, x, labels, mask = _inputs(g)
    n_train = int(mask.sum())

    eager = _model_from_golden(g, mode)
    opt_e = torch.optim.Adam(_train_params(eager), lr=0.01, weight_decay=1e-3, capturable=True)
    graphed = _model_from_golden(g, mode)
    opt_g = torch.optim.Adam(_train_params(graphed), lr=0.01, weight_decay=1e-3, capturable=True)
    p0 = [p.detach().clone() for p in graphed.parameters()]

    step = GraphedTrainStep(graphed, opt_g, x, (op, None, None), labels, mask, warmup=3)
    # the warm-up steps were rolled back: parameters untouched, optimizer state at step 0
    for p, q in zip(graphed.parameters(), p0):
        assert torch.equal(p.detach(), q)
    for st in opt_g.state.values():
        assert float(st["step"]) == 0.0 and float(st["exp_avg"].abs().max()) == 0.0
    assert step.launches_per_step >= 8

    eager.train()
    losses_e, losses_g = [], []
    for i in range(5):
        opt_e.zero_grad(set_to_none=True)
        loss = nll_log_softmax(eager(x, op, None, None), labels, mask, n_train=n_train)
        loss.backward()
        if i == 0:
            grads_e = {k: p.grad.detach().clone() for k, p in eager.named_parameters() if p.grad is not None}
        opt_e.step()
        losses_e.append(float(loss))
        losses_g.append(float(step()))
        if i == 0:
            # first replay: gradients of the captured step == eager gradients (atomics reorder sums: tiny noise)
            for k, p in graphed.named_parameters():
                if k in grads_e:
                    ref = grads_e[k]
                    scale = float(ref.abs().max()) + 1e-12
                    assert float((p.grad - ref).abs().max()) <= 2e-4 * scale, k
    assert abs(losses_e[0] - losses_g[0]) <= 1e-5 * abs(losses_e[0])
    # Adam amplifies sign flips of ~zero gradient entries; the trajectories still agree closely
    for a, b in zip(losses_e, losses_g):
        assert abs(a - b) <= 2e-2 * abs(a), (losses_e, losses_g)
    assert losses_g[-1] < losses_g[0]
    # the oracle value of the very first loss (same parameters, same inputs)
    tol = 1e-4 if mode == "fp32" else 3e-2
    assert abs(losses_g[0] - float(g.z["loss"])) <= tol * abs(float(g.z["loss"]))


def test_graphed_forward_matches_eager_and_tracks_parameter_updates():
    from acm_gnn_b200.graphed import GraphedForward
    g = Golden("gcn_pt_acmgcn_v0")
    op, x, labels, mask = _inputs(g)
    model = _model_from_golden(g, "fp32")
    fwd = GraphedForward(model, x, (op, None, None))
    model.eval()
    with torch.no_grad():
        ref = model(x, op, None, None)
    assert torch.equal(fwd(), ref)
    assert float((fwd().cpu() - torch.from_numpy(g.z["out"])).abs().max()) <= 2e-5 * float(abs(g.z["out"]).max())
    # parameters are read by the replay, not baked in
    with torch.no_grad():
        model.gcns[1].weight_low.mul_(0.5)
        ref2 = model(x, op, None, None)
    assert torch.equal(fwd(), ref2)
    assert not torch.equal(ref, ref2)


def test_graphed_step_with_reference_style_adjacency_and_new_split():
    """Adjacency passed the way ACM-Pytorch/train.py passes it (dense adj_low, COO adj_high); a
    new split of the same size is written into the static mask in place."""
    from acm_gnn_b200.graphed import GraphedTrainStep
    g = Golden("gcn_pt_acmgcn_v0")
    low, high, un = g.adjacency()
    low, high = low.cuda(), high.cuda()
    _, x, labels, mask = _inputs(g)
    model = _model_from_golden(g, "fp32")
    opt = torch.optim.Adam(_train_params(model), lr=0.01, capturable=True)
    step = GraphedTrainStep(model, opt, x, (low, high, None), labels, mask)
    l0 = float(step())
    assert abs(l0 - float(g.z["loss"])) <= 1e-4 * abs(float(g.z["loss"]))
    perm = torch.randperm(g.n, device="cuda")
    step.train_mask.copy_(mask[perm])      # same |train|, different rows
    l1 = float(step())
    assert l1 == l1 and l1 != l0


def test_graphed_step_with_dropout_trains():
    import acm_gnn_b200 as A
    from acm_gnn_b200.graphed import GraphedTrainStep
    os.environ["ACMB200_DTYPE"] = "bf16"
    g = Golden("gcn_pt_acmgcn_v0")
    op, x, labels, mask = _inputs(g)
    torch.manual_seed(0)
    model = A.GCN(g.nfeat, 64, g.nclass, 2, g.n, 0.5, "acmgcn", 0, variant=False).cuda()
    opt = torch.optim.Adam(_train_params(model), lr=0.05, capturable=True)
    step = GraphedTrainStep(model, opt, x, (op, None, None), labels, mask)
    losses = [float(step()) for _ in range(25)]
    assert all(l == l for l in losses) and losses[-1] < losses[0]


def test_graphed_step_rejects_bad_setups():
    import acm_gnn_b200 as A
    from acm_gnn_b200.graphed import GraphedTrainStep
    g = Golden("gcn_pt_acmgcn_v0")
    op, x, labels, mask = _inputs(g)
    model = _model_from_golden(g, "fp32")
    opt = torch.optim.Adam(_train_params(model), lr=0.01)          # capturable=False
    with pytest.raises(ValueError, match="capturable"):
        GraphedTrainStep(model, opt, x, (op, None, None), labels, mask)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        GraphedTrainStep(model, opt, x.cpu(), (op, None, None), labels, mask)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("f", [5, 16, 30])
def test_narrow_row_hint_bit_identical(f, mode):
    """acm_set_narrow_row_hint only changes the L2 prefetch-size qualifier of the neighbour-row
    loads of the narrow-row gathers (fused forward and transposed backward): same bits out."""
    import acm_gnn_b200 as A
    from acm_gnn_b200 import _lib
    os.environ["ACMB200_DTYPE"] = mode
    n = 3000
    row, col = O.synthetic_edges(n, 50000, seed=7)
    op = A.AcmOperator.from_edges(torch.from_numpy(row).cuda(), torch.from_numpy(col).cuda(), n)
    torch.manual_seed(1)
    layer = A.GraphConvolution(40, f, n, "acmgcn", variant=False).cuda()
    x0 = torch.rand(n, 40, device="cuda")
    res = []
    try:
        for hint in (0, 1):
            _lib.call("acm_set_narrow_row_hint", hint)
            x = x0.clone().requires_grad_(True)      # input gradient -> transform-first order + transposed gather
            layer.zero_grad(set_to_none=True)
            y = layer(x, op, None, None)
            y.square().sum().backward()
            res.append((y.detach().clone(), x.grad.detach().clone()))
    finally:
        _lib.call("acm_set_narrow_row_hint", 1)     # the default
    assert torch.isfinite(res[0][0]).all()
    assert torch.equal(res[0][0], res[1][0])
    assert torch.equal(res[0][1], res[1][1])


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("f", [5, 16, 64])
def test_degree_sorted_row_order_bit_identical(f, mode, monkeypatch):
    """The gather kernels may process the rows in a degree-sorted order (CsrMatrix.row_order):
    the per-row accumulation order is unchanged, so outputs and gradients are bitwise equal to the
    natural order.  Skewed degrees (1 .. > 256: long rows included) and a row count that is not a
    multiple of the rows per CTA."""
    import acm_gnn_b200 as A
    os.environ["ACMB200_DTYPE"] = mode
    n = 5003
    g = torch.Generator().manual_seed(11)
    src = torch.randint(0, n, (60000,), generator=g)
    dst = (n * torch.rand(60000, generator=g, dtype=torch.float64).pow(3.0)).long().clamp(max=n - 1)
    row, col = torch.cat([src, dst]).cuda(), torch.cat([dst, src]).cuda()
    torch.manual_seed(1)
    layer = A.GraphConvolution(40, f, n, "acmgcn", variant=False).cuda()
    x0 = torch.rand(n, 40, device="cuda")
    res = []
    for on in ("1", "0"):
        monkeypatch.setenv("ACMB200_ROW_ORDER", on)
        op = A.AcmOperator.from_edges(row, col, n)          # the order is cached per operator
        order = op.low.row_order()
        if on == "1":
            assert order is not None and torch.equal(torch.sort(order.long()).values, torch.arange(n, device="cuda"))
            deg = (op.low.rowptr[1:] - op.low.rowptr[:-1])[order.long()]
            w = op.low.ORDER_WINDOW
            assert bool((deg[:w][1:] >= deg[:w][:-1]).all()) and int(deg.max()) > 256
            assert int(order[:w].max()) < w                  # windows keep the rows local
        else:
            assert order is None
        x = x0.clone().requires_grad_(True)
        layer.zero_grad(set_to_none=True)
        y = layer(x, op, None, None)
        y.square().sum().backward()
        res.append((y.detach().clone(), x.grad.detach().clone(), layer.att_low.detach().clone()))
    assert torch.isfinite(res[0][0]).all()
    # rows with > 256 edges go through the segment-parallel pass, whose fp32 atomics make even two
    # identical runs differ in the last bits; every other row must be bitwise equal
    deg = op.low.rowptr[1:] - op.low.rowptr[:-1]
    short = deg <= op.low.LONG_ROW
    assert int((~short).sum()) > 0
    for a, b in zip(res[0][::2], res[1][::2]):                 # y and att are per-row quantities
        assert torch.equal(a[short], b[short])
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6)
    assert torch.allclose(res[0][1], res[1][1], rtol=1e-3, atol=1e-6)   # dX mixes rows through A^T


def test_graphed_step_with_staged_input_update_and_operator_lifetime():
    """(i) A StagedInput inside a captured step: new feature values reach the captured kernels only through
    ``StagedInput.update_`` (which refreshes the kernel-layout copy in place).  (ii) The captured graph bakes the CSR
    pointers of the operator in; the Graphed* classes resolve the operator once and keep it alive even when the
    conversion cache evicts its own reference."""
    import acm_gnn_b200 as A
    from acm_gnn_b200 import operator as opmod
    from acm_gnn_b200.functional import nll_log_softmax
    from acm_gnn_b200.graphed import GraphedForward, GraphedTrainStep
    g = Golden("gcn_pt_acmgcn_v0")
    mode = "bf16"
    model = _model_from_golden(g, mode)
    opt = torch.optim.Adam(_train_params(model), lr=0.01, capturable=True)
    low, high, un = g.adjacency()                      # the reference's own tensors (dense adj_low, COO adj_high)
    low, high = low.cuda(), high.cuda()
    x, labels = g.x.cuda(), g.labels.cuda()
    mask = torch.zeros(g.n, dtype=torch.uint8, device="cuda")
    mask[g.idx_train.cuda()] = 1
    staged = A.stage_input(x.clone(), mode)
    step = GraphedTrainStep(model, opt, staged, (low, high, None), labels, mask, warmup=2)
    assert isinstance(step.adj[0], A.AcmOperator)      # resolved once, strong reference
    opmod._CACHE.clear()                               # the cache forgets it; the graph must not care
    junk = [torch.empty(1 << 20, device="cuda") for _ in range(8)]   # churn the allocator
    del junk
    l0 = float(step())
    # eager reference of the same first step
    ref = _model_from_golden(g, mode)
    opt_r = torch.optim.Adam(_train_params(ref), lr=0.01, capturable=True)
    opt_r.zero_grad(set_to_none=True)
    loss_r = nll_log_softmax(ref(x, step.adj[0], None, None), labels, mask)
    assert abs(l0 - float(loss_r)) <= 1e-3 * abs(float(loss_r)) + 1e-5
    # new features: only update_ refreshes what the captured kernels read
    fwd = GraphedForward(model, staged, (low, high, None))
    out_a = fwd().clone()
    x2 = torch.rand_like(x)
    x2 = x2 / x2.sum(1, keepdim=True)
    staged.update_(x2)
    out_b = fwd().clone()
    assert not torch.equal(out_a, out_b)
    model.eval()
    with torch.no_grad():
        out_ref = model(A.stage_input(x2, mode), step.adj[0], None, None)
    assert torch.allclose(out_b, out_ref, rtol=1e-3, atol=1e-4)
