"""Justification of the bf16-mode gradient tolerance (tests/test_gpu_parity.py): rounding the
SAME storage points the CUDA path keeps in bf16 (layer input, weights, [HL|HH|HI] table,
dS / dH gradient tables) inside the fp32 oracle already moves parameter gradients by several
per cent of max|ref| on a small graph, while the forward output stays within 1e-2."""
import torch
import torch.nn.functional as F

from helpers import O


def _q(t):  # value rounded to bf16, gradient passed through
    return t + (t.bfloat16().float() - t).detach()


class _QGrad(torch.autograd.Function):  # identity whose incoming gradient is rounded to bf16
    @staticmethod
    def forward(ctx, a):
        return a

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def _layer(p, x, low, emul):
    ws = [p[k] for k in ("weight_low", "weight_high", "weight_mlp")]
    if emul:
        x, ws = _q(x), [_q(w) for w in ws]
    hl, hh, hi = [x @ w for w in ws]
    if emul:
        hl, hh, hi = [_QGrad.apply(_q(t)) for t in (hl, hh, hi)]
    sl, sh = torch.sparse.mm(low, hl), hh - torch.sparse.mm(low, hh)
    if emul:
        sl, sh = _QGrad.apply(sl), _QGrad.apply(sh)
    ol, oh, oi = F.relu(sl), F.relu(sh), F.relu(hi)
    z = torch.cat([ol @ p["att_vec_low"], oh @ p["att_vec_high"], oi @ p["att_vec_mlp"]], 1)
    att = torch.softmax(torch.sigmoid(z) @ p["att_vec"] / 3, 1)
    return 3 * (att[:, 0:1] * ol + att[:, 1:2] * oh + att[:, 2:3] * oi)


def test_bf16_storage_noise_level():
    torch.manual_seed(0)
    n, fin, hid, ncls = 600, 32, 64, 7
    row, col = O.synthetic_edges(n, 6000, seed=0)
    low, _ = O.operator_to_torch(O.build_operator(row, col, n))
    x = O.row_normalise_features(torch.rand(n, fin))
    labels, idx = torch.randint(0, ncls, (n,)), torch.randperm(n)[:360]
    base = O.init_gcn_params(fin, hid, ncls, n, "acmgcn", 0, torch.Generator().manual_seed(42))

    def run(emul):
        p = {g: {k: v.clone().requires_grad_(True) for k, v in d.items()} for g, d in base.items()}
        out = _layer(p["gcns.1"], F.relu(_layer(p["gcns.0"], x, low, emul)), low, emul)
        O.train_step_loss(out, labels, idx).backward()
        return out.detach(), p

    o0, p0 = run(False)
    o1, p1 = run(True)
    fwd = float((o1 - o0).abs().max() / o0.abs().max())
    assert fwd < 1e-2
    worst = 0.0
    for k in ("weight_low", "weight_high", "weight_mlp", "att_vec_low", "att_vec_high", "att_vec_mlp"):
        a, b = p1["gcns.0"][k].grad, p0["gcns.0"][k].grad
        worst = max(worst, float((a - b).norm() / b.norm()))
    # the noise is real (well above fp32 round-off) and of the magnitude the GPU tolerance allows
    assert 5e-3 < worst < 0.35, worst
