"""Parity AT THE HEADLINE SHAPE: Fin = hidden = 256, 16 classes, mean degree 20 -- the widths, degree and
model of the bench workload (SURVEY 8d cfg 5) on the N = 100 k / E = 2 M sample that bench.py's CPU arm
times (``cpu_step_factory``) -- instead of only on few-hundred-node graphs.  The same inputs and parameters go
through the GPU model (fp32 and bf16 storage, aggregate-first and transform-first order, staged and raw
input) and through the CPU oracle (oracle/acm_oracle.py, pinned to the reference by the golden vectors);
output, loss, attention columns and EVERY gradient are compared.

The oracle runs twice: in fp32 (the reference's arithmetic) and in fp64 ("truth").  Sums over 10^5 rows make
the fp32 oracle itself inexact (the high-pass weight gradient is a cancelling sum: measured 1e-4..4e-4
relative), so gradients are judged against the fp64 run.

Stated tolerances at this size:
  fp32 storage : output / attention max err <= 2e-5 * max|ref| vs the fp32 oracle; every gradient within
                 relative Frobenius 3e-3 of the fp64 truth (or 4x the fp32 oracle's own distance, whichever is
                 larger): the parameter gradients are cancelling sums over 10^5 rows that the kernels
                 accumulate in fp32 in a different order (per-thread partial sums, then atomics) than
                 ATen's blocked reductions -- measured 3e-7 .. 3e-4, and 1.2e-3 for the high-pass weight
                 gradient, whose summands (dS_H - A^T dS_H on smooth signals) cancel to ~1e-3 of their size;
                 the reference's own fp32 run is 2.7e-4 away from fp64 on that tensor;
  bf16 storage : forward max err <= 3e-2 * max|ref| as everywhere.  Gradients: the round-1 review expected the
                 bf16 storage noise to average out at 10^5 nodes (<= 2e-2); MEASURED it does not -- relative
                 Frobenius 6e-2 .. 1.4e-1 (cosine 0.994 .. 0.9999), about the same as on the small golden
                 graphs.  The reason is cancellation, not the kernels: with all-positive, row-normalised
                 features and softmax-CE gradients that sum to ~0 over the nodes, dW = sum_i x_i (x) dh_i
                 cancels to a few per cent of its summands, so the 2^-9 relative rounding of the STORED
                 dh_i / S_i rows survives as a several-per-cent error of the sum, at any N.  The bound is
                 therefore tied to an emulation: the fp32 oracle with the same storage points rounded to
                 bf16 (tests/test_bf16_noise_cpu.py::_layer) run on THIS workload gives the noise level of
                 bf16 storage itself; every GPU gradient must be within 4x that level (and <= 0.25 relative
                 Frobenius, cosine >= 0.99).  Kernel logic is what the fp32 mode checks at 3e-3 above.
"""
import os

import numpy as np
import pytest
import torch

from helpers import O

pytestmark = pytest.mark.gpu

N, E, FIN, HID, NCLS = 100_000, 2_000_000, 256, 256, 16
BF16_FRO, BF16_COS = 0.25, 0.99
FP32_FRO = 3e-3


@pytest.fixture(scope="module")
def workload():
    """Inputs + the oracle's results (CPU, fp32, torch.sparse.mm COO -- the reference's own op sequence)."""
    row, col = O.synthetic_edges(N, E, seed=0)
    op_ref = O.build_operator(row, col, N, "pytorch")
    low, high = O.operator_to_torch(op_ref)
    g = torch.Generator().manual_seed(1)
    x = O.row_normalise_features(torch.rand(N, FIN, generator=g))
    labels = torch.randint(0, NCLS, (N,), generator=g)
    idx = torch.randperm(N, generator=g)[: int(0.6 * N)]
    gp = torch.Generator().manual_seed(42)
    params = O.init_gcn_params(FIN, HID, NCLS, 0, "acmgcn", 0, gp)
    torch.set_num_threads(os.cpu_count())

    def run(dt):
        ps = {grp: {k: t.detach().to(dt).clone() for k, t in d.items()} for grp, d in params.items()}
        for grp in ps.values():
            for k, t in grp.items():
                if not k.startswith(("layer_norm", "struc", "att_struc")):
                    t.requires_grad_(True)
        out, atts = O.gcn_forward(ps, x.to(dt), low.to(dt), high.to(dt), None)
        loss = O.train_step_loss(out, labels, idx)
        loss.backward()
        return {"out": out.detach(), "loss": float(loss), "atts": [a.detach() for a in atts],
                "grads": {f"{grp}.{k}": t.grad.clone() for grp, d in ps.items() for k, t in d.items() if t.grad is not None}}

    ref, ref64 = run(torch.float32), run(torch.float64)

    # noise level of bf16 STORAGE on this workload: fp32 oracle with the storage points rounded to bf16
    from test_bf16_noise_cpu import _layer as q_layer
    ps = {grp: {k: t.detach().clone() for k, t in d.items()} for grp, d in params.items()}
    for grp in ps.values():
        for k, t in grp.items():
            if not k.startswith(("layer_norm", "struc", "att_struc")):
                t.requires_grad_(True)
    out_q = q_layer(ps["gcns.1"], torch.relu(q_layer(ps["gcns.0"], x, low, True)), low, True)
    O.train_step_loss(out_q, labels, idx).backward()
    noise = {f"{grp}.{k}": float((t.grad.double() - ref64["grads"][f"{grp}.{k}"]).norm() / ref64["grads"][f"{grp}.{k}"].norm())
             for grp, d in ps.items() for k, t in d.items() if t.grad is not None}
    sd = {f"{grp}.{k}": t.detach().clone() for grp, d in params.items() for k, t in d.items()}
    return dict(row=row, col=col, op_ref=op_ref, x=x, labels=labels, idx=idx, ref=ref, ref64=ref64, sd=sd, noise=noise)


def _rel(got, ref):
    got = got.detach().double().cpu().ravel()
    ref = ref.detach().double().cpu().ravel()
    fro = float((got - ref).norm() / ref.norm().clamp_min(1e-300))
    cos = float((got @ ref) / (got.norm() * ref.norm()).clamp_min(1e-300))
    mx = float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-300))
    return fro, cos, mx


@pytest.mark.parametrize("staged", [False, True])
@pytest.mark.parametrize("order", ["auto", "off"])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_headline_shape_matches_oracle(workload, mode, order, staged, monkeypatch):
    import acm_gnn_b200 as A
    from acm_gnn_b200 import _lib
    from acm_gnn_b200.functional import nll_log_softmax
    w = workload
    monkeypatch.setenv("ACMB200_DTYPE", mode)
    monkeypatch.setenv("ACMB200_REORDER", order)
    op = A.AcmOperator.from_edges(torch.from_numpy(w["row"]).cuda(), torch.from_numpy(w["col"]).cuda(), N)
    assert np.array_equal(op.low.rowptr.cpu().numpy(), w["op_ref"].rowptr)
    assert np.array_equal(op.low.val.cpu().numpy().view(np.uint32), w["op_ref"].w_low.view(np.uint32))
    model = A.GCN(FIN, HID, NCLS, 2, N, 0.0, "acmgcn", 0, variant=False).cuda()
    missing, unexpected = model.load_state_dict({k: v for k, v in w["sd"].items() if not k.endswith("struc_low")}, strict=False)
    assert not unexpected, unexpected
    x = w["x"].cuda()
    xin = A.stage_input(x, mode) if staged else x
    mask = torch.zeros(N, dtype=torch.uint8, device="cuda")
    mask[w["idx"].cuda()] = 1
    timer = _lib.KernelTimer()
    _lib.set_timer(timer)
    try:
        model.train()
        out = model(xin, op, None, None)
        loss = nll_log_softmax(out, w["labels"].cuda(), mask)
        loss.backward()
        torch.cuda.synchronize()
    finally:
        _lib.set_timer(None)
    names = set(k.split(":")[0] for k in timer.spans)
    assert ("acm_spmm_agg_first" in names), names          # order auto: forward; order off: input-side backward
    ref = w["ref"]
    tol = 2e-5 if mode == "fp32" else 3e-2
    fro, cos, mx = _rel(out, ref["out"])
    assert mx <= tol, f"out: max rel err {mx:.3e} ({mode}, order {order}, staged {staged})"
    assert abs(float(loss) - ref["loss"]) <= (2e-5 if mode == "fp32" else 2e-3) * abs(ref["loss"]), (float(loss), ref["loss"])
    for li, layer in enumerate(model.gcns):
        att = torch.cat([layer.att_low, layer.att_high, layer.att_mlp], 1)
        assert _rel(att, ref["atts"][li])[2] <= tol, f"att{li}"
    report, bad = {}, []
    for k, p in model.named_parameters():
        if k not in ref["grads"]:
            continue
        assert p.grad is not None, k
        truth = w["ref64"]["grads"][k]
        fro, cos, mx = _rel(p.grad, truth)
        fro_ref = _rel(ref["grads"][k], truth)[0]          # how inexact the reference's own fp32 arithmetic is
        report[k] = (fro, cos, fro_ref)
        if mode == "fp32":
            ok = fro <= max(4 * fro_ref, FP32_FRO)
        else:
            ok = fro <= min(BF16_FRO, max(4 * w["noise"][k], 2e-2)) and cos >= BF16_COS
        if not ok:
            bad.append(k)
    table = "; ".join(f"{k}: {v[0]:.2e} (cos {v[1]:.6f}, fp32 oracle {v[2]:.1e}, bf16-storage emulation {w['noise'].get(k, float('nan')):.1e})"
                      for k, v in sorted(report.items()))
    assert not bad, f"gradients out of tolerance ({mode}, order {order}, staged {staged}): {bad}\nrel.fro vs fp64 -- {table}"
    assert len(report) >= 14, sorted(report)
    worst = max(report.items(), key=lambda kv: kv[1][0])
    print(f"headline shape [{mode}, order {order}, staged {staged}]: out max rel err {_rel(out, ref['out'])[2]:.2e}, "
          f"worst gradient {worst[0]} rel.fro {worst[1][0]:.2e} cos {worst[1][1]:.6f} (fp32 oracle {worst[1][2]:.1e}) | {table}")
