"""Inter-layer glue kernel (acm_glue_fwd / acm_glue_bwd) against the reference's own ops
``F.dropout(F.relu(fea1), p, training) + xX`` (ACM-Pytorch/models/models.py:160-164):

  * dropout inactive (the default-fused case): output AND gradients BIT-IDENTICAL to the torch ops, fp32 / bf16 /
    bf16 + fp32 ``xX`` (torch's type promotion), ragged sizes (tail of < 8 elements), with and without relu / add;
  * dropout active (opt-in): the pass bits equal oracle/philox_oracle.py's numpy restatement of
    Philox4x32-10 BIT FOR BIT (through the C ABI and through autograd), y == relu(x) * mask / (1 - p) + add in the
    reference's rounding order, dx == g * mask / (1 - p); the device-resident offset advances by one per launch and a
    captured CUDA graph draws a fresh mask on every replay.
"""
import numpy as np
import pytest
import torch

from helpers import ROOT  # noqa: F401
from oracle import philox_oracle as P

pytestmark = pytest.mark.gpu

SHAPES = [(1001, 37), (256, 64), (3, 1), (5, 8)]


def _ref_ops(x, add, relu, p, training):
    y = torch.nn.functional.dropout(torch.relu(x) if relu else x, p, training=training)
    return y if add is None else y + add


@pytest.mark.parametrize("dtypes", [("fp32", "fp32"), ("bf16", "bf16"), ("bf16", "fp32")])
@pytest.mark.parametrize("relu", [False, True])
@pytest.mark.parametrize("with_add", [False, True])
def test_glue_without_dropout_is_bit_identical_to_torch(dtypes, relu, with_add):
    from acm_gnn_b200 import _lib
    from acm_gnn_b200.functional import inter_layer_glue
    dt = {"fp32": torch.float32, "bf16": torch.bfloat16}
    xdt, adt = dt[dtypes[0]], dt[dtypes[1]]
    if not with_add and xdt != adt:
        pytest.skip("the mixed case needs the add")
    if not relu and not with_add:
        pytest.skip("identity: no launch")
    g = torch.Generator(device="cuda").manual_seed(3)
    for shape in SHAPES:
        x = torch.randn(shape, device="cuda", generator=g).to(xdt)
        add = torch.randn(shape, device="cuda", generator=g).to(adt) if with_add else None
        for training in (True, False):
            xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
            aa = add.clone().requires_grad_(True) if with_add else None
            ab = add.clone().requires_grad_(True) if with_add else None
            n0 = _lib.launch_count()
            ya = inter_layer_glue(xa, aa, relu, 0.0 if training else 0.5, training)
            assert _lib.launch_count() == n0 + 1, "the fused launch did not run"
            yb = _ref_ops(xb, ab, relu, 0.0 if training else 0.5, training)
            assert ya.dtype == yb.dtype and torch.equal(ya, yb), (shape, training)
            go = torch.randn(shape, device="cuda", generator=g).to(ya.dtype)
            ya.backward(go)
            yb.backward(go)
            assert xa.grad.dtype == xb.grad.dtype and torch.equal(xa.grad, xb.grad), (shape, training)
            if with_add:
                assert torch.equal(aa.grad, ab.grad)
    # no gradient requested: no mask is written
    with torch.no_grad():
        assert torch.equal(inter_layer_glue(x, add, relu, 0.0, True), _ref_ops(x, add, relu, 0.0, True))


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_glue_dropout_mask_matches_philox_oracle_bit_for_bit(mode, monkeypatch):
    from acm_gnn_b200 import _lib
    from acm_gnn_b200 import functional as Fn
    from acm_gnn_b200.functional import _stream, glue_rng_state, inter_layer_glue
    monkeypatch.setenv("ACMB200_FUSED_DROPOUT", "1")
    Fn._GLUE_RNG.clear()                                    # independent of what ran before
    dt = torch.float32 if mode == "fp32" else torch.bfloat16
    p = 0.3
    torch.manual_seed(20261017)
    state = glue_rng_state(torch.device("cuda", 0))
    seed, off0 = (int(v) for v in state.cpu())
    assert seed == torch.cuda.default_generators[0].initial_seed() and off0 == 0
    g = torch.Generator(device="cuda").manual_seed(4)
    for k, shape in enumerate([(1001, 37), (4096, 256), (7, 3)]):
        total = shape[0] * shape[1]
        x = torch.randn(shape, device="cuda", generator=g).to(dt)
        add = torch.randn(shape, device="cuda", generator=g).to(dt)
        # (a) straight through the C ABI: the mask bytes
        y = torch.empty_like(x)
        mask = torch.zeros((total + 7) // 8, dtype=torch.uint8, device="cuda")
        _lib.call("acm_glue_fwd", int(dt == torch.bfloat16), int(dt == torch.bfloat16), x.data_ptr(), add.data_ptr(), y.data_ptr(),
                  mask.data_ptr(), total, 1, p, state.data_ptr(), _stream())
        off = off0 + 2 * k
        y_ref, on_ref = P.glue_forward(x.cpu(), add.cpu(), True, p, seed, off)
        assert np.array_equal(mask.cpu().numpy(), P.pack_mask(on_ref.numpy().ravel())), shape
        assert torch.equal(y.cpu(), y_ref), shape
        assert int(state[1]) == off + 1
        # (b) through autograd: next offset, value and gradient
        xa = x.clone().requires_grad_(True)
        aa = add.clone().requires_grad_(True)
        ya = inter_layer_glue(xa, aa, True, p, True)
        y_ref, on_ref = P.glue_forward(x.cpu(), add.cpu(), True, p, seed, off + 1)
        assert torch.equal(ya.detach().cpu(), y_ref)
        go = torch.randn(shape, device="cuda", generator=g).to(dt)
        ya.backward(go)
        scale = float(np.float32(1.0) / (np.float32(1.0) - np.float32(p)))
        dx_ref = torch.where(on_ref, (go.cpu().float() * scale).to(dt), torch.zeros((), dtype=dt))
        assert torch.equal(xa.grad.cpu(), dx_ref) and torch.equal(aa.grad, go)
        keep = P.keep_bits(seed, off + 1, total, p)
        if total > 10000:
            assert abs(keep.mean() - (1 - p)) < 5e-3
            # dropout semantics of the reference op: E[y] = relu(x) (+ add)
            err = float((ya.detach().float() - add.float()).mean() - torch.relu(x.float()).mean())
            assert abs(err) < 2e-2
    # eval: the generator is not consulted
    before = int(state[1])
    assert torch.equal(inter_layer_glue(x, None, True, p, False), torch.relu(x))
    assert int(state[1]) == before
    # a new torch seed re-seeds the glue's generator
    torch.manual_seed(99)
    s2 = glue_rng_state(torch.device("cuda", 0))
    assert int(s2[0]) == 99 and int(s2[1]) == 0


def test_glue_dropout_in_a_cuda_graph_draws_fresh_masks(monkeypatch):
    from acm_gnn_b200.functional import glue_rng_state, inter_layer_glue
    from acm_gnn_b200 import functional as Fn
    monkeypatch.setenv("ACMB200_FUSED_DROPOUT", "1")
    Fn._GLUE_RNG.clear()
    torch.manual_seed(11)
    x = torch.rand(2048, 64, device="cuda") + 0.5
    state = glue_rng_state(x.device)
    inter_layer_glue(x, None, True, 0.5, True)                # eager warm-up (offset 0 -> 1)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        inter_layer_glue(x, None, True, 0.5, True)
        torch.cuda.current_stream().synchronize()
        with torch.cuda.graph(graph, stream=s):
            y = inter_layer_glue(x, None, True, 0.5, True)
    torch.cuda.current_stream().wait_stream(s)
    off = int(state[1])
    outs = []
    for i in range(3):
        graph.replay()
        torch.cuda.synchronize()
        outs.append(y.clone())
        assert int(state[1]) == off + i + 1
        _, on = P.glue_forward(x.cpu(), None, True, 0.5, 11, off + i)
        assert torch.equal(outs[-1].cpu() != 0, on)
    assert not torch.equal(outs[0], outs[1]) and not torch.equal(outs[1], outs[2])


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_model_with_fused_glue_matches_model_with_torch_glue(mode, monkeypatch):
    """acmgcn++ (xX branch), variant 1 (live relu): the stack with the fused glue is bit-identical to the stack with
    the reference's torch ops between the layers: output bit for bit, gradients to the run-to-run noise of the fp32
    atomics that reduce the parameter gradients (relative Frobenius 1e-5)."""
    import acm_gnn_b200 as A
    from helpers import O
    monkeypatch.setenv("ACMB200_DTYPE", mode)
    n, e, fin, hid, ncls = 3000, 30000, 96, 64, 5
    row, col = O.synthetic_edges(n, e, seed=2)
    op = A.AcmOperator.from_edges(torch.from_numpy(row).cuda(), torch.from_numpy(col).cuda(), n)
    g = torch.Generator().manual_seed(0)
    x = O.row_normalise_features(torch.rand(n, fin, generator=g)).cuda()
    labels = torch.randint(0, ncls, (n,), generator=g).cuda()
    res = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("ACMB200_FUSED_GLUE", fused)
        torch.manual_seed(7)
        model = A.GCN(fin, hid, ncls, 2, n, 0.0, "acmgcnpp", 0, variant=True).cuda()
        model.train()
        out = model(x, op, None, None)
        loss = torch.nn.functional.nll_loss(torch.log_softmax(out, 1), labels)
        loss.backward()
        res[fused] = (out.detach(), {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None})
    assert torch.equal(res["1"][0], res["0"][0])
    assert res["1"][1].keys() == res["0"][1].keys() and len(res["1"][1]) >= 14
    for k in res["1"][1]:
        a, b = res["1"][1][k].double(), res["0"][1][k].double()
        assert float((a - b).norm()) <= 1e-5 * float(b.norm()) + 1e-12, k
