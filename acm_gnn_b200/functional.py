"""Autograd boundary of the ACM layer: one ``torch.autograd.Function`` per layer call whose
forward and backward are sequences of libacm_b200 launches on the current CUDA stream.

Math (validated against the reference, SURVEY.md section 8a / DESIGN.md):
  fwd  [HL|HH|HI] = X [W_L|W_H|W_I]
       variant 0: O_L = relu(A HL), O_H = relu(HH - A HH)        (layers.py:188-193)
       variant 1: O_L = A relu(HL), O_H = relu(HH) - A relu(HH)  (layers.py:178-184)
       O_I = relu(HI) ; [O_S = relu(A_raw S)] ; att = softmax(sigmoid([LN](O_k).a_k) Avec / K)
       Y = c * sum_k att_k O_k       (c = 3, or 1 with the structure channel; layers.py:200-232)
  bwd  mix_bwd (row local) -> transposed aggregation -> dWcat = X^T dH, dX = dH Wcat^T
Aggregate-first order (variant 0, input without gradient, pad(Fin) <= 2*FP; default):
  fwd  Z = A X ; D = X - Z ; [S_L|S_H|HI] = [Z W_L | D W_H | X W_I] ; epilogue as above
  bwd  dW_L = Z^T dS_L ; dW_H = D^T dS_H ; dW_I = X^T dHI          (no transposed aggregation)
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from .operator import AcmOperator

_FP_CHOICES = (8, 16, 32, 64, 128, 256)


def padded_width(f: int) -> int:
    for c in _FP_CHOICES:
        if f <= c:
            return c
    raise NotImplementedError(f"out_features={f} > 256 is not supported by the fused ACM kernels yet")


def _stream():
    return torch.cuda.current_stream().cuda_stream


@dataclass
class LayerConfig:
    variant: bool = False
    k_channels: int = 3
    ln_live: bool = False
    out_scale: float = 3.0
    dtype: str = "bf16"        # storage of feature tables: "bf16" | "fp32" (accumulation is fp32)
    gemm: str = "auto"         # "simt" | "tcgen05" | "auto"
    reorder: str = "env"       # aggregate-first order A(XW) = (AX)W: "auto" | "off" | "env" (ACMB200_REORDER)
    dist: Optional[object] = None   # acm_gnn_b200.dist.RowPartition or None
    layer_key: int = 0              # identifies the layer's persistent symmetric-memory tables
    out_dtype: str = "fp32"         # dtype of Y: "fp32" (reference boundary) | "bf16" (inter-layer activations)

    def storage(self):
        return (torch.bfloat16, _lib.ACM_BF16) if self.dtype == "bf16" else (torch.float32, _lib.ACM_F32)

    def gemm_impl(self, fin: int):
        g = self.gemm
        if g == "auto":
            g = os.environ.get("ACMB200_GEMM", "auto")
        if g == "auto":
            g = "tcgen05" if (self.dtype == "bf16" and tc_available()) else "simt"
        if g == "tcgen05" and self.dtype != "bf16":
            raise ValueError("the tcgen05 GEMM path computes in bf16; use dtype='bf16' or gemm='simt'")
        return _lib.GEMM_TCGEN05 if g == "tcgen05" else _lib.GEMM_SIMT


_TC_OK = None


def tc_available() -> bool:
    """tcgen05 GEMM kernels (gemm_tc.cu) need an sm_100 device (the library is built for sm_100a
    only) -- they are the default for bf16 storage there; ACMB200_GEMM=simt forces the CUDA-core path."""
    global _TC_OK
    if _TC_OK is None:
        _TC_OK = torch.cuda.is_available() and torch.cuda.get_device_capability()[0] == 10
    return _TC_OK


def reorder_enabled(cfg: LayerConfig) -> bool:
    r = cfg.reorder
    if r == "env":
        r = os.environ.get("ACMB200_REORDER", "auto").lower()
    if r in ("auto", "1", "on"):
        return True
    if r in ("off", "0"):
        return False
    raise ValueError(f"ACMB200_REORDER={r!r}: expected auto or off")


def use_aggregate_first(cfg: LayerConfig, fin: int, fp: int, x_needs_grad: bool) -> bool:
    """SURVEY 8(f) rank 4.  A(XW_L) = (AX)W_L and HH - A(XW_H) = (X - AX)W_H: aggregate the
    layer INPUT once (Fin wide) instead of the [HL|HH] table (2*out_features wide).  Valid when
    the relu sits after the aggregation (variant 0); chosen when the input row is not wider
    than the table row and the input needs no gradient (first layer) -- then the backward
    needs no transposed aggregation at all: dW_L = (AX)^T dS_L, dW_H = (X-AX)^T dS_H."""
    if not reorder_enabled(cfg) or cfg.variant or x_needs_grad or fin > 256:
        return False
    return padded_width(fin) <= 2 * fp


def _knob(name: str, default: str = "auto") -> bool:
    v = os.environ.get(name, default).lower()
    if v in ("auto", "1", "on"):
        return True
    if v in ("off", "0"):
        return False
    raise ValueError(f"{name}={v!r}: expected auto or off")


def use_input_backward(cfg: LayerConfig, fin: int, fp: int, x_needs_grad: bool) -> bool:
    """Transform-first variant-0 layer whose input needs no gradient: the transposed aggregation
    only feeds the weight gradients, dW_L = X^T (A^T dS_L) = (A X)^T dS_L and
    dW_H = X^T (dS_H - A^T dS_H) = (X - A X)^T dS_H, so the backward aggregates the layer INPUT
    (Fin wide, recomputed every step -- nothing is cached across steps) instead of gathering the
    2*out_features wide [dS_L|dS_H] table.  Under a row partition the input rows of all ranks are
    already resident (StagedInput) or narrower to all-gather than the table, so the backward of
    such a layer needs NO exchange at all.  ``ACMB200_BWD_INPUT=off`` keeps the transposed
    aggregation of the reference's autograd order."""
    if cfg.variant or x_needs_grad or fin > 256 or not _knob("ACMB200_BWD_INPUT"):
        return False
    return padded_width(fin) <= 2 * fp


def use_rank1_table(cfg: LayerConfig, op, fp: int) -> bool:
    """Variant 1 (relu before the aggregation) without LayerNorm: the rows handed to the transposed
    aggregation are dO_k = c att_k G + dz_k a_k^T, so the backward gathers (and, under a row partition,
    exchanges) the G row plus four scalars per node instead of the 2*out_features wide [dO_L|dO_H] row --
    half the bytes (csrc/spmm_t.cu spmm_t_rank1_kernel).  Wide rows only (padded width >= 64) and only when
    the transposed operator has no long rows (those keep the segment-parallel pass of the plain table).
    ``ACMB200_BWD_RANK1`` = auto (default): under a row partition only -- there it halves the NVLink exchange;
    on one GPU the rank-1 gather issues twice the instructions per byte (18 FMAs + the scalar broadcast per
    512-byte row) and measured 38.5 ms against 35 ms for the plain table at the headline size; on: always; off."""
    v = os.environ.get("ACMB200_BWD_RANK1", "auto").lower()
    if v in ("off", "0"):
        return False
    if v not in ("auto", "on", "1"):
        raise ValueError(f"ACMB200_BWD_RANK1={v!r}: expected auto, on or off")
    if v == "auto" and cfg.dist is None:
        return False
    if not (bool(cfg.variant) and not cfg.ln_live and fp >= 64):
        return False
    return _no_long_rows_on_any_rank(op, cfg.dist)


def _no_long_rows_on_any_rank(op, part) -> bool:
    """The table layout is a protocol between the ranks (every rank writes rows into every other rank's table), so the
    choice must be the same everywhere: the rank-1 table is used only when NO rank's slice of the transposed operator
    has long rows.  Decided once per operator (one tiny all-reduce at its first backward)."""
    flag = getattr(op, "_no_long_t_anywhere", None)
    if flag is None:
        local = op.low.long_rows(True) is None
        if part is None:
            flag = local
        else:
            t = torch.tensor([0.0 if local else 1.0], device=op.low.col.device)
            part.all_reduce_(t)
            flag = float(t.item()) == 0.0
        op._no_long_t_anywhere = flag
    return flag


FUSED_FWD_DEFAULT = "auto"


def use_fused_forward(cfg: LayerConfig, impl: int, fp: int, f: int, k_channels: int, ldx: int) -> bool:
    """Aggregate-first layer at out_features (padded) = 256 in bf16 storage on the tcgen05 path: one launch
    computes [S_L|S_H|HI] = [Z W_L | D W_H | X W_I] into TMEM and applies the attention / mix epilogue from
    there (csrc/fused_fwd.cu) -- no [S_L|S_H|HI] round trip through HBM.  Three channels without LayerNorm
    only; y rows must be 16-byte aligned.  ``ACMB200_FUSED_FWD=off`` keeps the three GEMM launches plus the
    pre-aggregated epilogue launch."""
    return (impl == _lib.GEMM_TCGEN05 and cfg.dtype == "bf16" and fp == 256 and k_channels == 3 and not cfg.ln_live
            and not cfg.variant and ldx % 8 == 0 and f % (16 if cfg.out_dtype == "bf16" else 8) == 0
            and _knob("ACMB200_FUSED_FWD", FUSED_FWD_DEFAULT))


def use_local_table(cfg: LayerConfig, ldx: int, fp: int) -> bool:
    """Row partition, transform-first order (SURVEY 7 "what to all-gather"): exchange the NARROWER
    of {layer input X, [HL|HH] table}.  When the input row is not wider than the table row every
    rank builds the whole [HL|HH] table itself from the all-gathered input (one redundant tcgen05
    GEMM, ~3 ms at the headline size) instead of receiving (P-1)/P of a 2*out_features wide table
    over NVLink (10.24 GB per step at the headline size).  ``ACMB200_LOCAL_TABLE=off`` keeps the
    fused push / NCCL all-gather of the table."""
    return cfg.dist is not None and ldx <= 2 * fp and _knob("ACMB200_LOCAL_TABLE")


def default_dtype() -> str:
    """Storage of the feature tables behind the reference-facing modules.  The reference computes in
    fp32, so an unmodified train.py launched through run.py gets fp32 tables (the mode held to the
    2e-5 parity tolerance) unless the user opts in to bf16 storage with ACMB200_DTYPE=bf16 (what
    BASELINE configs 2-5 name, and what bench.py passes explicitly)."""
    d = os.environ.get("ACMB200_DTYPE", "fp32").lower()
    if d in ("bf16", "bfloat16"):
        return "bf16"
    if d in ("fp32", "f32", "float32"):
        return "fp32"
    raise ValueError(f"ACMB200_DTYPE={d!r}: expected bf16 or fp32")


def build_pack(fp, f, a_vecs, att_vec, ln_params):
    """Pack the channel-attention parameters in the layout of include/acm_b200.h."""
    dev = att_vec.device
    pack = torch.zeros(16 * fp + 32, dtype=torch.float32, device=dev)
    for k, a in enumerate(a_vecs):
        pack[k * fp:k * fp + f] = a.reshape(-1)
    kk = att_vec.shape[0]
    pack[4 * fp:4 * fp + 16].view(4, 4)[:kk, :kk] = att_vec
    if ln_params is not None:
        for k, (g, b) in enumerate(ln_params):
            a = a_vecs[k].reshape(-1)
            pack[4 * fp + 16 + k * fp:4 * fp + 16 + k * fp + f] = g
            pack[8 * fp + 16 + k * fp:8 * fp + 16 + k * fp + f] = b
            pack[12 * fp + 16 + k * fp:12 * fp + 16 + k * fp + f] = g * a       # derived tail
            pack[16 * fp + 16 + k] = (b * a).sum()
            pack[16 * fp + 20 + k] = (g * a).sum()
    return pack


def build_wcat(fp, f, ws, dtype):
    fin = ws[0].shape[0]
    wcat = torch.zeros(fin, 3 * fp, dtype=torch.float32, device=ws[0].device)
    for k, w in enumerate(ws):
        wcat[:, k * fp:k * fp + f] = w
    return wcat.to(dtype)


def _long_pass(csr, transposed, table, fp, halves, cdt, st):
    """Segment-parallel aggregation of the long rows of ``csr`` (degree skew).  Returns the
    (long_rows_ptr, n_long, acc_ptr, keepalive) arguments of the row kernels."""
    lr = csr.long_rows(transposed)
    if lr is None:
        return 0, 0, 0, None
    rows, seg_long, e0, e1 = lr
    col = csr.col_t if transposed else csr.col
    val = csr.val_t if transposed else csr.val
    acc = torch.zeros(rows.numel(), halves * fp, dtype=torch.float32, device=rows.device)
    _lib.call("acm_spmm_long_rows", cdt, fp, halves, seg_long.numel(), seg_long.data_ptr(), e0.data_ptr(), e1.data_ptr(),
              col.data_ptr(), val.data_ptr(), table.data_ptr(), acc.data_ptr(), st)
    return rows.data_ptr(), int(rows.numel()), acc.data_ptr(), (rows, acc)


def _row_order(csr, transposed, fp):
    """Degree-sorted processing order for the gather kernels when several rows share a warp
    (fp <= 64: 4 or more rows per warp); None = natural order."""
    return csr.row_order(transposed) if fp <= 64 else None


def _ptr_array(tensors):
    """HOST array of device pointers (NULL for None) for the `const float* const*` ABI arguments."""
    return (ctypes.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])


def pack_params(cfg, fp, f, fin, ldt, ws, a_vecs, att_vec, ln_params, need_wt, st):
    """wcat [fin,3fp] T, wcat_t [3fp,ldt] T (or None) and the attention pack in ONE launch."""
    tdt, cdt = cfg.storage()
    dev = att_vec.device
    ws = [w.detach().contiguous() for w in ws]
    a_vecs = [a.detach().contiguous() for a in a_vecs]
    av = att_vec.detach().contiguous()
    wcat = torch.empty(fin, 3 * fp, dtype=tdt, device=dev)
    wcat_t = torch.empty(3 * fp, ldt, dtype=tdt, device=dev) if need_wt else None
    pack = torch.empty(16 * fp + 32, dtype=torch.float32, device=dev)  # ACM_PACK_FLOATS_VALUE
    k = len(a_vecs)
    a_arr = _ptr_array(a_vecs)
    g_arr = b_arr = None
    keep = [ws, a_vecs, av]
    if ln_params is not None:
        gs = [g.detach().contiguous() for g, _ in ln_params]
        bs = [b.detach().contiguous() for _, b in ln_params]
        g_arr, b_arr = _ptr_array(gs), _ptr_array(bs)
        keep += [gs, bs]
    _lib.call("acm_pack_params", cdt, fin, f, fp, k, int(ln_params is not None), ldt,
              ws[0].data_ptr(), ws[1].data_ptr(), ws[2].data_ptr(), ctypes.addressof(a_arr), av.data_ptr(),
              0 if g_arr is None else ctypes.addressof(g_arr), 0 if b_arr is None else ctypes.addressof(b_arr),
              wcat.data_ptr(), _lib.ptr(wcat_t), pack.data_ptr(), st)
    del keep
    return wcat, wcat_t, pack


def _stage_rows(x, tdt, cdt, ldx, st):
    """Copy of the layer input in the storage dtype with row stride ``ldx`` (zero padded)."""
    n, fin = x.shape
    xc = x.detach().contiguous()
    if xc.dtype == tdt and ldx == fin:
        return xc
    if xc.dtype != torch.float32:
        xs = torch.zeros(n, ldx, dtype=tdt, device=x.device)
        xs[:, :fin] = xc
        return xs
    xs = torch.empty(n, ldx, dtype=tdt, device=x.device)
    _lib.call("acm_cast_pad", xc.data_ptr(), n, fin, fin, xs.data_ptr(), cdt, ldx, st)
    return xs


def _aggregate_input(op, x_all, n, ldx, tdt, cdt, st):
    """Z = A_low X and D = X - Z for the own rows (one gather of the INPUT row per stored edge)."""
    z = torch.empty(n, ldx, dtype=tdt, device=x_all.device)
    d = torch.empty(n, ldx, dtype=tdt, device=x_all.device)
    lr = _long_pass(op.low, False, x_all, ldx, 1, cdt, st)
    _lib.call("acm_spmm_agg_first", cdt, ldx, n, op.row0, op.low.rowptr.data_ptr(), op.low.col.data_ptr(),
              op.low.val.data_ptr(), x_all.data_ptr(), z.data_ptr(), d.data_ptr(), lr[0], lr[1], lr[2], st, tag=ldx)
    return z, d


class StagedInput:
    """Layer-0 input features already resident in HBM in the layout the kernels consume: the
    storage-dtype copy padded to the power-of-two width (``xs`` [n_local, ldx]) and, under a row
    partition, the all-gathered table of every rank's rows (``x_all`` [N, ldx]).

    The reference keeps its fp32 features resident on the device for the whole run and passes
    the same tensor every epoch (ACM-Pytorch/utils.py:383-385); staging is the analogue for the
    bf16 / partitioned layout: done once at data-placement time (``stage_input``), it removes
    the per-step cast and -- multi-GPU -- the per-step all-gather of static input data.  A plain
    fp32 tensor is always accepted too (reference API; staged on every call)."""

    def __init__(self, x, xs, x_all, dtype, dist=None):
        self.x, self.xs, self.x_all, self.dtype, self.dist = x, xs, x_all, dtype, dist
        self.shape = x.shape
        self.device = x.device
        self.is_cuda = x.is_cuda

    def update_(self, x_new: torch.Tensor) -> "StagedInput":
        """New feature values INTO THE SAME BUFFERS (the only supported way to change a staged input:
        captured CUDA graphs and saved pointers keep reading ``xs`` / ``x_all``).  Re-runs the cast/pad
        and, under a row partition, the all-gather."""
        if x_new.shape != self.x.shape:
            raise ValueError(f"update_: expected shape {tuple(self.x.shape)}, got {tuple(x_new.shape)}")
        if self.x.data_ptr() != x_new.data_ptr():
            self.x.copy_(x_new)
        n, fin = self.x.shape
        if self.xs.data_ptr() != self.x.data_ptr():
            cdt = _lib.ACM_BF16 if self.dtype == "bf16" else _lib.ACM_F32
            xc = self.x.detach().contiguous()
            _lib.call("acm_cast_pad", xc.data_ptr(), n, fin, fin, self.xs.data_ptr(), cdt, self.xs.shape[1], _stream())
        if self.dist is not None:
            self.x_all.copy_(self.dist.all_gather_rows(self.xs))
        return self


def stage_input(x: torch.Tensor, dtype: Optional[str] = None, dist=None) -> StagedInput:
    if not x.is_cuda:
        raise RuntimeError("acm_gnn_b200: CUDA tensors only (there is no CPU fallback)")
    dtype = dtype or default_dtype()
    n, fin = x.shape
    if fin > 256:
        raise NotImplementedError("stage_input: in_features > 256 (the aggregate-first order does not apply)")
    tdt, cdt = (torch.bfloat16, _lib.ACM_BF16) if dtype == "bf16" else (torch.float32, _lib.ACM_F32)
    ldx = padded_width(fin)
    xc = x.detach().contiguous()
    if dtype == "fp32" and ldx == fin:
        xs = xc
    else:
        xs = torch.empty(n, ldx, dtype=tdt, device=x.device)
        _lib.call("acm_cast_pad", xc.data_ptr(), n, fin, fin, xs.data_ptr(), cdt, ldx, _stream())
    x_all = xs if dist is None else dist.all_gather_rows(xs)
    return StagedInput(x, xs, x_all, dtype, dist)


class AcmLayerFunction(torch.autograd.Function):
    """Inputs: op, cfg, x, W_low, W_high, W_mlp, a_low, a_high, a_mlp, att_vec, struc_low,
    a_struc, then (only when LayerNorm is live) gamma_k, beta_k per channel.
    Outputs: Y [n, F] (differentiable), att [n, K] (non-differentiable)."""

    @staticmethod
    def forward(ctx, op: AcmOperator, cfg: LayerConfig, x, w_low, w_high, w_mlp, a_low, a_high, a_mlp, att_vec,
                struc_low, a_struc, *ln_flat):
        if not x.is_cuda:
            raise RuntimeError("acm_gnn_b200: CUDA tensors only (there is no CPU fallback)")
        staged = x if isinstance(x, StagedInput) else None
        if staged is not None:
            if staged.dtype != cfg.dtype:
                raise ValueError(f"input staged as {staged.dtype} but the layer runs in {cfg.dtype}")
            x = staged.x
        tdt, cdt = cfg.storage()
        n, fin = x.shape
        f = w_low.shape[1]
        fp = padded_width(f)
        K = cfg.k_channels
        dev = x.device
        st = _stream()
        if n != op.n:
            raise ValueError(f"input has {n} rows but the operator has {op.n}")

        a_vecs = [a_low, a_high, a_mlp] + ([a_struc] if K == 4 else [])
        ln_params = None
        if cfg.ln_live:
            ln_params = [(ln_flat[2 * k], ln_flat[2 * k + 1]) for k in range(K)]
        impl = cfg.gemm_impl(fin)
        x_needs_grad = bool(ctx.needs_input_grad[2])
        # grad mode is off inside Function.forward: needs_input_grad is the reliable signal
        # (all False under torch.no_grad(), e.g. ACM-Geometric's evaluate_acmgcn)
        need_grad = any(ctx.needs_input_grad)
        agg_first = use_aggregate_first(cfg, fin, fp, x_needs_grad)
        bwd_input = (not agg_first) and use_input_backward(cfg, fin, fp, x_needs_grad)
        # row stride of the staged input = K extent of the transposed weights
        if agg_first or bwd_input:
            ldx = padded_width(fin)          # the input-row gather wants a power-of-two table width
        elif staged is not None:
            ldx = staged.xs.shape[1]
        else:
            ldx = (fin + 7) // 8 * 8 if cfg.dtype == "bf16" else fin
        wcat, wcat_t_all, pack = pack_params(cfg, fp, f, fin, ldx, (w_low, w_high, w_mlp), a_vecs, att_vec, ln_params,
                                             impl == _lib.GEMM_TCGEN05, st)
        # staging copy of the layer input in the storage dtype, [n, ldx]; under a row partition x_all
        # holds the rows of every rank (resident in a StagedInput, all-gathered on demand otherwise)
        if staged is not None:
            if staged.xs.shape[1] != ldx:
                raise ValueError("StagedInput was staged for a different layer width")
            xs, x_all = staged.xs, staged.x_all
        else:
            if x.dtype == torch.bfloat16 and cfg.dtype != "bf16":
                raise ValueError("bf16 input needs ACMB200_DTYPE=bf16")
            xs, x_all = _stage_rows(x, tdt, cdt, ldx, st), None
        h_i = torch.empty(n, fp, dtype=tdt, device=dev)
        z = d = wcat_t = h_lh = None
        fused = False
        if agg_first:
            # ---- aggregate-first: Z = A X, D = X - Z, then [S_L|S_H|HI] = [Z W_L | D W_H | X W_I] ----
            fused = use_fused_forward(cfg, impl, fp, f, K, padded_width(fin))
            # the fused kernel keeps [S_L|S_H] in TMEM: the table only exists in HBM when the backward needs it
            h_lh = torch.empty(n, 2 * fp, dtype=tdt, device=dev) if (need_grad or not fused) else None
            if x_all is None:
                x_all = xs if cfg.dist is None else cfg.dist.all_gather_rows(xs)
            z, d = _aggregate_input(op, x_all, n, ldx, tdt, cdt, st)
            x_all = None
            if ldx == fin:
                wp = wcat
            else:
                wp = torch.zeros(ldx, 3 * fp, dtype=tdt, device=dev)   # rows fin..ldx are zero
                wp[:fin] = wcat
            wt = wcat_t_all  # [3fp, ldx], K-major (tcgen05 path) or None
            if not fused:
                for k, (a_op, c_ptr, ldc) in enumerate(((z, h_lh.data_ptr(), 2 * fp),
                                                        (d, h_lh[:, fp:].data_ptr(), 2 * fp),
                                                        (xs, h_i.data_ptr(), fp))):
                    _lib.call("acm_gemm_ab", impl, cdt, a_op.data_ptr(), ldx, wp[:, k * fp:].data_ptr(), 3 * fp,
                              0 if wt is None else wt[k * fp:].data_ptr(), ldx, c_ptr, ldc, n, fp, ldx, 0, st, tag=fp)
            del wp
            table, csr = h_lh, (0, 0, 0)      # pre-aggregated: only the epilogue of the fused kernel runs
            row0 = 0
            lr = (0, 0, 0, None)
            order = None
        else:
            wcat_t = wcat_t_all
            local_tab = use_local_table(cfg, ldx, fp)
            push = (cfg.dist is not None and not local_tab and impl == _lib.GEMM_TCGEN05 and cfg.dist.push_enabled())
            if local_tab:
                # exchange the narrower operand: every rank builds the WHOLE [HL|HH] table from the
                # input rows of all ranks (redundant GEMM) -- no table crosses NVLink
                if x_all is None:
                    x_all = cfg.dist.all_gather_rows(xs)
                table = torch.empty(x_all.shape[0], 2 * fp, dtype=tdt, device=dev)
                _lib.call("acm_gemm_ab", impl, cdt, x_all.data_ptr(), ldx, wcat.data_ptr(), 3 * fp,
                          _lib.ptr(wcat_t), ldx, table.data_ptr(), 2 * fp, x_all.shape[0], 2 * fp, fin,
                          int(cfg.variant), st, tag=f"table{fp}")
                _lib.call("acm_gemm_ab", impl, cdt, xs.data_ptr(), ldx, wcat[:, 2 * fp:].data_ptr(), 3 * fp,
                          0 if wcat_t is None else wcat_t[2 * fp:].data_ptr(), ldx, h_i.data_ptr(), fp, n, fp, fin,
                          0, st, tag=fp)
                h_lh = table[op.row0:op.row0 + n]
            elif push:
                # fused GEMM + all-gather: the epilogue stores every finished [HL|HH] row into all
                # ranks' tables through NVLink peer mappings; one device-side barrier afterwards
                # (tables alternate between two buffers, see RowPartition.symm_table)
                table, hdl, ptrs, mc = cfg.dist.symm_table((cfg.layer_key, "fwd"), 2 * fp, tdt, dev)
                _lib.call("acm_gemm_xw_fwd_push", xs.data_ptr(), ldx, wcat_t.data_ptr(), ctypes.addressof(ptrs),
                          cfg.dist.world, op.row0, mc, h_i.data_ptr(), n, fin, fp, int(cfg.variant), st, tag=fp)
                hdl.barrier(channel=0)     # every rank's rows have landed everywhere
                h_lh = table[op.row0:op.row0 + n]
            else:
                h_lh = torch.empty(n, 2 * fp, dtype=tdt, device=dev)
                _lib.call("acm_gemm_xw_fwd", impl, cdt, xs.data_ptr(), ldx, wcat.data_ptr(), _lib.ptr(wcat_t),
                          h_lh.data_ptr(), h_i.data_ptr(), n, fin, fp, int(cfg.variant), st, tag=fp)
                table = h_lh if cfg.dist is None else cfg.dist.all_gather_rows(h_lh)
            if bwd_input:
                if x_all is None:   # the backward aggregates the input rows of every rank
                    x_all = xs if cfg.dist is None else cfg.dist.all_gather_rows(xs)
            else:
                x_all = None
            csr = (op.low.rowptr.data_ptr(), op.low.col.data_ptr(), op.low.val.data_ptr())
            row0 = op.row0
            lr = _long_pass(op.low, False, table, fp, 2, cdt, st)
            order = _row_order(op.low, False, fp)

        o_s = None
        if K == 4:
            if op.raw is None:
                raise ValueError("structure_info=1 needs adj_low_unnormalized")
            s_loc = struc_low.detach().contiguous()
            s_tab = torch.empty(s_loc.shape[0], fp, dtype=tdt, device=dev)
            _lib.call("acm_cast_pad", s_loc.data_ptr(), s_loc.shape[0], f, f, s_tab.data_ptr(), cdt, fp, st)
            o_s = torch.empty(n, fp, dtype=tdt, device=dev)
            _lib.call("acm_spmm_plain", cdt, cdt, fp, n, op.raw.rowptr.data_ptr(), op.raw.col.data_ptr(),
                      op.raw.val.data_ptr(), s_tab.data_ptr(), o_s.data_ptr(), fp, fp, 1, st)
            del s_tab

        y_bf16 = (cfg.out_dtype == "bf16")
        if y_bf16 and cfg.dtype != "bf16":
            raise ValueError("bf16 activations need ACMB200_DTYPE=bf16")
        y = torch.empty(n, f, dtype=torch.bfloat16 if y_bf16 else torch.float32, device=dev)
        att = torch.empty(n, K, dtype=torch.float32, device=dev)
        sig = torch.empty(n, K, dtype=torch.float32, device=dev) if need_grad else None
        # aggregate-first: the table already holds [S_L|S_H] and O_k = relu(S_k), so the table itself is
        # what the backward needs (mix_bwd applies the relu on load for variant 0) -- no second copy
        o_save = torch.empty(n, 2 * fp, dtype=tdt, device=dev) if (need_grad and not agg_first) else None
        if fused:
            # the three GEMMs and the attention / mix epilogue in ONE tcgen05 launch: the fp32 accumulators
            # stay in TMEM, [S_L|S_H] and HI are written once (bf16) for the backward -- or not at all
            _lib.call("acm_fused_agg_fwd", z.data_ptr(), d.data_ptr(), xs.data_ptr(), ldx, wt.data_ptr(), ldx, pack.data_ptr(),
                      n, ldx, f, fp, float(cfg.out_scale), y.data_ptr(), _lib.ACM_BF16 if y_bf16 else _lib.ACM_F32, f,
                      _lib.ptr(h_lh), h_i.data_ptr(), att.data_ptr(), _lib.ptr(sig),
                      st, tag=fp)
        else:
            _lib.call("acm_spmm_mix_fwd", cdt, fp, f, n, row0, csr[0], csr[1], csr[2], 0,
                      table.data_ptr(), h_i.data_ptr(), _lib.ptr(o_s), pack.data_ptr(),
                      K, int(cfg.ln_live), int(cfg.variant), float(cfg.out_scale),
                      y.data_ptr(), _lib.ACM_BF16 if y_bf16 else _lib.ACM_F32, f, _lib.ptr(o_save), att.data_ptr(), _lib.ptr(sig),
                      lr[0], lr[1], lr[2], _lib.ptr(order), st, tag=fp)
        del lr
        if need_grad and agg_first:
            o_save = h_lh

        ctx.mark_non_differentiable(att)
        if need_grad:
            ctx.op, ctx.cfg = op, cfg
            ctx.dims = (n, fin, f, fp, K, ldx, impl)
            ctx.agg_first = agg_first
            ctx.bwd_input = bwd_input
            ctx.x_all = x_all if bwd_input else None     # resident (staged) or all-gathered input rows
            ctx.x_dtype = x.dtype
            ctx.x_needs_grad = x_needs_grad
            ctx.n_ln = len(ln_flat)
            ctx.struc_rows = 0 if struc_low is None else struc_low.shape[0]
            # variant 1 needs the relu'd forward table (its positivity is the relu mask)
            ctx.save_for_backward(xs, wcat, wcat_t, h_i, o_save, att, sig, pack, o_s,
                                  h_lh if cfg.variant else None, z, d)
        return y, att

    @staticmethod
    def backward(ctx, g, _g_att):
        xs, wcat, wcat_t, h_i, o_save, att, sig, pack, o_s, p_tab, z, d = ctx.saved_tensors
        op, cfg = ctx.op, ctx.cfg
        n, fin, f, fp, K, ldx, impl = ctx.dims
        tdt, cdt = cfg.storage()
        dev = g.device
        st = _stream()
        g = g.contiguous()
        if g.dtype not in (torch.float32, torch.bfloat16):
            g = g.to(torch.float32)
        gdt = _lib.ACM_BF16 if g.dtype == torch.bfloat16 else _lib.ACM_F32

        dh_all = torch.empty(n, 3 * fp, dtype=tdt, device=dev)
        dos_pre = torch.empty(n, fp, dtype=tdt, device=dev) if K == 4 else None
        # parameter-gradient accumulators of this layer in ONE buffer: [dwcat (fin x 3fp) | dpack]  -> one memset and,
        # under a row partition, one all-reduce instead of two
        gbuf = torch.zeros(fin * 3 * fp + 12 * fp + 16, dtype=torch.float32, device=dev)
        dpack = gbuf[fin * 3 * fp:]
        needs_t = not (ctx.agg_first or ctx.bwd_input)      # a transposed aggregation follows
        rank1 = needs_t and use_rank1_table(cfg, op, fp)
        tw = (fp + (8 if cfg.dtype == "bf16" else 4)) if rank1 else 2 * fp     # table bytes per row / element size
        t_rows = n if cfg.dist is None else cfg.dist.world * cfg.dist.rows_per_rank
        push = (cfg.dist is not None and needs_t and cfg.dist.push_enabled())
        if push:
            # fused mix_bwd + all-gather of the backward operand table (peer stores over NVLink)
            t_table, hdl, ptrs, mc = cfg.dist.symm_table((cfg.layer_key, "bwd"), tw, tdt, dev)
            _lib.call("acm_mix_bwd", cdt, fp, f, n, g.data_ptr(), gdt, f, o_save.data_ptr(), h_i.data_ptr(), _lib.ptr(o_s),
                      att.data_ptr(), sig.data_ptr(), pack.data_ptr(), K, int(cfg.ln_live), int(cfg.variant),
                      float(cfg.out_scale), 0, int(rank1), t_rows, dh_all.data_ptr(), _lib.ptr(dos_pre), dpack.data_ptr(),
                      ctypes.addressof(ptrs), cfg.dist.world, op.row0, mc, st, tag=fp)
            hdl.barrier(channel=0)
            t_lh = None
        else:
            t_lh = torch.empty(n, tw, dtype=tdt, device=dev)
            _lib.call("acm_mix_bwd", cdt, fp, f, n, g.data_ptr(), gdt, f, o_save.data_ptr(), h_i.data_ptr(), _lib.ptr(o_s),
                      att.data_ptr(), sig.data_ptr(), pack.data_ptr(), K, int(cfg.ln_live), int(cfg.variant),
                      float(cfg.out_scale), t_lh.data_ptr(), int(rank1), n, dh_all.data_ptr(), _lib.ptr(dos_pre), dpack.data_ptr(),
                      0, 0, 0, 0, st, tag=fp)

        dwcat = gbuf[:fin * 3 * fp].view(fin, 3 * fp)
        dx = None
        if ctx.bwd_input:
            # transform-first forward, input without gradient: aggregate the input rows here instead of
            # gathering the [dS_L|dS_H] table (use_input_backward) -- no exchange under a partition
            z, d = _aggregate_input(op, ctx.x_all, n, ldx, tdt, cdt, st)
            ctx.x_all = None
        if ctx.agg_first or ctx.bwd_input:
            # dW_L = (AX)^T dS_L ; dW_H = (X - AX)^T dS_H ; dW_I = X^T dHI -- no transposed aggregation
            for k, (a_op, b_ptr, ldb) in enumerate(((z, t_lh.data_ptr(), 2 * fp),
                                                    (d, t_lh[:, fp:].data_ptr(), 2 * fp),
                                                    (xs, dh_all[:, 2 * fp:].data_ptr(), 3 * fp))):
                _lib.call("acm_gemm_atb", impl, cdt, a_op.data_ptr(), ldx, b_ptr, ldb,
                          dwcat[:, k * fp:].data_ptr(), 3 * fp, n, fin, fp, st, tag=fp)
        else:
            if not push:
                if cfg.dist is None:
                    t_table = t_lh
                elif rank1:
                    # NCCL path: the rank-1 table is two regions (G rows, then the scalars) -> gather each
                    flat = t_lh.view(-1)
                    g_all = cfg.dist.all_gather_rows(flat[:n * fp].view(n, fp))
                    s_all = cfg.dist.all_gather_rows(flat[n * fp:].view(n, tw - fp))
                    t_table = torch.empty(g_all.shape[0], tw, dtype=tdt, device=dev)
                    t_table.view(-1)[:g_all.numel()].copy_(g_all.view(-1))
                    t_table.view(-1)[g_all.numel():].copy_(s_all.view(-1))
                    del flat, g_all, s_all
                else:
                    t_table = cfg.dist.all_gather_rows(t_lh)
            if rank1:
                _lib.call("acm_spmm_t_bwd_rank1", cdt, fp, n, op.row0, op.low.rowptr_t.data_ptr(), op.low.col_t.data_ptr(),
                          op.low.val_t.data_ptr(), t_table.data_ptr(), t_table.shape[0], pack.data_ptr(), _lib.ptr(p_tab),
                          dh_all.data_ptr(), st, tag=fp)
                lr = None
            else:
                lr = _long_pass(op.low, True, t_table, fp, 2, cdt, st)
                _lib.call("acm_spmm_t_bwd", cdt, fp, n, op.row0, op.low.rowptr_t.data_ptr(), op.low.col_t.data_ptr(),
                          op.low.val_t.data_ptr(), t_table.data_ptr(), _lib.ptr(p_tab), dh_all.data_ptr(),
                          lr[0], lr[1], lr[2], _lib.ptr(_row_order(op.low, True, fp)), st, tag=fp)
            del t_table, lr
            _lib.call("acm_gemm_bwd_dw", impl, cdt, xs.data_ptr(), ldx, dh_all.data_ptr(), dwcat.data_ptr(), n, fin, fp, st, tag=fp)
            if ctx.x_needs_grad and ctx.x_dtype == torch.bfloat16 and fin % 8 == 0:
                # bf16 activations: the gradient goes back in bf16 as well (dX = dH Wcat^T)
                dx = torch.empty(n, fin, dtype=torch.bfloat16, device=dev)
                wt = wcat_t if wcat_t is not None else wcat.t().contiguous()
                _lib.call("acm_gemm_ab", impl, cdt, dh_all.data_ptr(), 3 * fp, wt.data_ptr(), wt.shape[1],
                          wcat.data_ptr(), 3 * fp, dx.data_ptr(), fin, n, fin, 3 * fp, 0, st, tag=fp)
            elif ctx.x_needs_grad:
                dx = torch.empty(n, fin, dtype=torch.float32, device=dev)
                _lib.call("acm_gemm_bwd_dx", impl, cdt, dh_all.data_ptr(), wcat.data_ptr(), _lib.ptr(wcat_t), ldx,
                          dx.data_ptr(), fin, n, fin, fp, st, tag=fp)
                if ctx.x_dtype != torch.float32:
                    dx = dx.to(ctx.x_dtype)
        del t_lh

        d_struc = d_a_struc = None
        if K == 4:
            nr = ctx.struc_rows
            if cfg.dist is None:
                d_struc = torch.empty(nr, f, dtype=torch.float32, device=dev)
                _lib.call("acm_spmm_plain", cdt, _lib.ACM_F32, fp, nr, op.raw.rowptr_t.data_ptr(), op.raw.col_t.data_ptr(),
                          op.raw.val_t.data_ptr(), dos_pre.data_ptr(), d_struc.data_ptr(), f, f, 0, st)
            else:
                # row partition: gather every rank's rows of d O_S, aggregate the own rows of
                # A_raw^T, then assemble the gradient of the REPLICATED struc_low parameter
                dos_all = cfg.dist.all_gather_rows(dos_pre)
                d_loc = torch.empty(n, f, dtype=torch.float32, device=dev)
                _lib.call("acm_spmm_plain", cdt, _lib.ACM_F32, fp, n, op.raw.rowptr_t.data_ptr(), op.raw.col_t.data_ptr(),
                          op.raw.val_t.data_ptr(), dos_all.data_ptr(), d_loc.data_ptr(), f, f, 0, st)
                d_struc = torch.zeros(nr, f, dtype=torch.float32, device=dev)
                d_struc[op.row0:op.row0 + n] = d_loc
                cfg.dist.all_reduce_(d_struc)
                del dos_all, d_loc

        if cfg.dist is not None:
            cfg.dist.all_reduce_(gbuf)

        # one launch splits dwcat / dpack into the reference's parameter shapes
        dw = [torch.empty(fin, f, dtype=torch.float32, device=dev) for _ in range(3)]
        da = [torch.empty(f, 1, dtype=torch.float32, device=dev) for _ in range(K)]
        d_att_vec = torch.empty(K, K, dtype=torch.float32, device=dev)
        n_ln_ch = ctx.n_ln // 2 if cfg.ln_live else 0
        dg = [torch.empty(f, dtype=torch.float32, device=dev) for _ in range(n_ln_ch)]
        db = [torch.empty(f, dtype=torch.float32, device=dev) for _ in range(n_ln_ch)]
        da_arr = _ptr_array(da)
        dg_arr = _ptr_array(dg + [None] * (K - n_ln_ch)) if n_ln_ch else None
        db_arr = _ptr_array(db + [None] * (K - n_ln_ch)) if n_ln_ch else None
        _lib.call("acm_unpack_grads", fin, f, fp, K, int(n_ln_ch > 0), dwcat.data_ptr(), dpack.data_ptr(),
                  dw[0].data_ptr(), dw[1].data_ptr(), dw[2].data_ptr(), ctypes.addressof(da_arr), d_att_vec.data_ptr(),
                  0 if dg_arr is None else ctypes.addressof(dg_arr), 0 if db_arr is None else ctypes.addressof(db_arr), st)
        if K == 4:
            d_a_struc = da[3]
        d_ln = []
        for k in range(ctx.n_ln // 2):
            if k < n_ln_ch:
                d_ln += [dg[k], db[k]]
            else:
                d_ln += [None, None]
        return (None, None, dx, dw[0], dw[1], dw[2], da[0], da[1], da[2], d_att_vec, d_struc, d_a_struc, *d_ln)


class MaskedNllLogSoftmax(torch.autograd.Function):
    """``F.nll_loss(F.log_softmax(out, 1)[train], labels[train])`` (ACM-Pytorch/utils.py:567-568)
    as one launch producing the loss and d loss / d out together (acm_nll_log_softmax)."""

    @staticmethod
    def forward(ctx, out, labels, mask, scale):
        if not out.is_cuda:
            raise RuntimeError("acm_gnn_b200: CUDA tensors only (there is no CPU fallback)")
        out_c = out.detach().contiguous()
        n, c = out_c.shape
        loss = torch.zeros((), dtype=torch.float32, device=out.device)
        need = ctx.needs_input_grad[0]
        d = torch.empty_like(out_c) if need else None
        _lib.call("acm_nll_log_softmax", out_c.data_ptr(), c, n, c, labels.data_ptr(), _lib.ptr(mask), float(scale),
                  loss.data_ptr(), _lib.ptr(d), c, _stream())
        if need:
            ctx.save_for_backward(d)
        return loss

    @staticmethod
    def backward(ctx, g):
        (d,) = ctx.saved_tensors
        return d * g, None, None, None


def nll_log_softmax(out, labels, train_mask=None, n_train=None):
    """Mean NLL of log_softmax(out) over the rows selected by ``train_mask`` (uint8/bool [N]).
    ``n_train`` overrides the normaliser (global count under a row partition).  The fused kernel
    takes fp32 CUDA logits with at most 64 classes; anything else goes through the reference's own
    torch glue (F.log_softmax + F.nll_loss, utils.py:567-568) -- same math, not a CPU fallback of
    the ACM layer.  Labels of the selected rows must lie in [0, C) (the kernel ignores rows whose
    label is out of range instead of reading out of bounds)."""
    if not out.is_cuda:
        raise RuntimeError("acm_gnn_b200: CUDA tensors only (there is no CPU fallback)")
    if out.dim() != 2:
        raise ValueError("nll_log_softmax: logits must be [N, C]")
    labels = labels.to(device=out.device, dtype=torch.int64).contiguous()
    if labels.shape[0] != out.shape[0]:
        raise ValueError("nll_log_softmax: labels must have one entry per row of the logits")
    if train_mask is not None:
        train_mask = train_mask.to(device=out.device, dtype=torch.uint8).contiguous()
        if train_mask.shape[0] != out.shape[0]:
            raise ValueError("nll_log_softmax: train_mask must have one entry per row of the logits")
    if n_train is None:
        n_train = int(train_mask.sum().item()) if train_mask is not None else out.shape[0]
    if n_train <= 0:
        raise ValueError("nll_log_softmax: no training rows selected")
    if out.dtype != torch.float32 or out.shape[1] > 64:
        lp = torch.nn.functional.log_softmax(out.float(), dim=1)
        sel = slice(None) if train_mask is None else train_mask.bool()
        return torch.nn.functional.nll_loss(lp[sel], labels[sel], reduction="sum") / float(n_train)
    return MaskedNllLogSoftmax.apply(out, labels, train_mask, 1.0 / float(n_train))


# ---- inter-layer glue: relu -> dropout -> (+ xX) between the two layers ----------
_GLUE_RNG = {}      # device index -> (torch seed it was derived from, int64[2] device tensor {seed, offset})


def glue_rng_state(device):
    """Device-resident {seed, offset} of the glue's Philox generator.  Seeded from torch's CUDA seed of
    that device (so ``torch.manual_seed`` controls it) and re-seeded (offset 0) whenever that seed CHANGES
    -- seeding torch again with the same value does not rewind it; the offset is advanced on the device by every dropout launch (graph replays draw fresh masks)."""
    torch.cuda.init()                                   # runs a pending lazy torch.manual_seed
    idx = device.index if device.index is not None else torch.cuda.current_device()
    seed = int(torch.cuda.default_generators[idx].initial_seed())
    cur = _GLUE_RNG.get(idx)
    if cur is None or cur[0] != seed:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("inter_layer_glue: run one eager step before capturing (the generator state is created lazily)")
        s = seed if seed < (1 << 63) else seed - (1 << 64)
        cur = (seed, torch.tensor([s, 0], dtype=torch.int64, device=torch.device("cuda", idx)))
        _GLUE_RNG[idx] = cur
    return cur[1]


class InterLayerGlue(torch.autograd.Function):
    """y = dropout(relu(x), p) [+ add]: one launch forward (acm_glue_fwd: value + one pass bit per
    element), one backward (acm_glue_bwd).  Replaces F.relu / F.dropout / "+" of
    ACM-Pytorch/models/models.py:160-164."""

    @staticmethod
    def forward(ctx, x, add, relu, p):
        xc = x.detach().contiguous()
        ac = add.detach().contiguous() if add is not None else None
        total = xc.numel()
        need = ctx.needs_input_grad[0]
        ydt = xc.dtype if ac is None else ac.dtype        # bf16 x + fp32 add -> fp32 (torch's promotion)
        y = torch.empty(xc.shape, dtype=ydt, device=x.device)
        mask = torch.empty((total + 7) // 8, dtype=torch.uint8, device=x.device) if need and (relu or p > 0) else None
        state = glue_rng_state(x.device) if p > 0 else None
        _lib.call("acm_glue_fwd", _dt_code(xc.dtype), _dt_code(ydt), xc.data_ptr(), _lib.ptr(ac), y.data_ptr(), _lib.ptr(mask),
                  total, int(bool(relu)), float(p), _lib.ptr(state), _stream())
        # the fp32 expression the kernel scaled by: 1.f / (1.f - p)
        pf = ctypes.c_float(p).value
        ctx.mask, ctx.xdt = mask, xc.dtype
        ctx.scale = ctypes.c_float(1.0 / ctypes.c_float(1.0 - pf).value).value if p > 0 else 1.0
        return y

    @staticmethod
    def backward(ctx, g):
        dx = None
        if ctx.needs_input_grad[0]:
            if ctx.mask is None:
                dx = g.to(ctx.xdt)
            else:
                gc = g.contiguous()
                dx = torch.empty(gc.shape, dtype=ctx.xdt, device=gc.device)
                _lib.call("acm_glue_bwd", _dt_code(ctx.xdt), _dt_code(gc.dtype), gc.data_ptr(), ctx.mask.data_ptr(), dx.data_ptr(),
                          gc.numel(), ctx.scale, _stream())
        return dx, (g if ctx.needs_input_grad[1] else None), None, None


def _dt_code(dt):
    return _lib.ACM_BF16 if dt == torch.bfloat16 else _lib.ACM_F32


def inter_layer_glue(x, add=None, relu=True, p=0.0, training=True):
    """``F.dropout(F.relu(x), p, training) + add`` -- the reference's glue between its two layers
    (ACM-Pytorch/models/models.py:160-164, ACM-Geometric/models.py:70-74; ``relu=False`` when the relu
    is the identity, ``add`` = the acmgcn++ ``xX`` branch).

    With the dropout inactive (eval, or p == 0) the fused launch is bit-identical to the torch ops and is
    the default.  With p > 0 in training the fused kernel draws its own Philox mask (same distribution,
    NOT torch's random stream), so by default the reference's three torch ops run and results stay
    reproducible against the reference under ``torch.manual_seed``; ``ACMB200_FUSED_DROPOUT=1`` opts in.
    ``ACMB200_FUSED_GLUE=0`` keeps the torch ops everywhere."""
    p_eff = float(p) if training and p >= 2.0 ** -32 else 0.0
    if not relu and add is None and p_eff == 0.0:
        return x
    fusable = (x.is_cuda and x.dtype in (torch.float32, torch.bfloat16) and 0.0 <= p_eff < 1.0
               and (add is None or (add.shape == x.shape and add.device == x.device
                                    and (add.dtype == x.dtype or (add.dtype == torch.float32 and x.dtype == torch.bfloat16))))
               and os.environ.get("ACMB200_FUSED_GLUE", "1") != "0"
               and (p_eff == 0.0 or os.environ.get("ACMB200_FUSED_DROPOUT", "0") == "1"))
    if not fusable:
        y = torch.nn.functional.dropout(torch.relu(x) if relu else x, p, training=training)
        return y if add is None else y + add
    return InterLayerGlue.apply(x, add, bool(relu), p_eff)


# ---- nn.Linear of the acmgcn++ mlpX branch on the tcgen05 path (bf16 storage mode) --------------------
class LinearBf16(torch.autograd.Function):
    """y = relu?(xs . W^T + b) in bf16 storage with fp32 accumulation (acm_linear_fwd), for an input that needs no
    gradient (the layer-0 features); dW = dY^T xs (acm_gemm_atb), db = column sums of dY."""

    @staticmethod
    def forward(ctx, xs, weight, bias, relu):
        m, ldx = xs.shape
        n, fin = weight.shape
        w = weight.detach()
        if ldx == fin:
            w_nk = w.to(torch.bfloat16).contiguous()
        else:
            w_nk = torch.zeros(n, ldx, dtype=torch.bfloat16, device=w.device)
            w_nk[:, :fin] = w
        b = bias.detach().float().contiguous() if bias is not None else None
        y = torch.empty(m, n, dtype=torch.bfloat16, device=xs.device)
        _lib.call("acm_linear_fwd", xs.data_ptr(), ldx, w_nk.data_ptr(), ldx, _lib.ptr(b), y.data_ptr(), n, m, n, ldx,
                  int(bool(relu)), _stream(), tag=f"{ldx}x{n}")
        ctx.save_for_backward(xs, y if relu else None)
        ctx.fin, ctx.relu, ctx.has_bias = fin, bool(relu), bias is not None
        return y

    @staticmethod
    def backward(ctx, g):
        xs, y = ctx.saved_tensors
        m, ldx = xs.shape
        d = g.to(torch.bfloat16).contiguous()
        if ctx.relu:
            d = torch.ops.aten.threshold_backward(d, y, 0)          # g where y > 0
        n = d.shape[1]
        dw = db = None
        if ctx.needs_input_grad[1]:
            dwp = torch.zeros(n, ldx, dtype=torch.float32, device=d.device)
            _lib.call("acm_gemm_atb", _lib.GEMM_TCGEN05, _lib.ACM_BF16, d.data_ptr(), n, xs.data_ptr(), ldx, dwp.data_ptr(), ldx,
                      m, n, ldx, _stream(), tag=f"linear{n}x{ldx}")
            dw = dwp if ldx == ctx.fin else dwp[:, :ctx.fin].contiguous()
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = d.sum(0, dtype=torch.float32)
        return None, dw, db, None


def linear_bf16_eligible(x, weight) -> bool:
    """The tcgen05 Linear applies to a CUDA input that needs no gradient (raw or staged layer-0 features) with
    out_features % 8 == 0 on an sm_100 device; everything else keeps torch's F.linear."""
    if not (weight.is_cuda and tc_available() and weight.shape[0] % 8 == 0 and weight.dtype == torch.float32):
        return False
    if isinstance(x, StagedInput):
        return x.dtype == "bf16" and x.xs.shape[1] >= weight.shape[1]
    return x.is_cuda and x.dim() == 2 and not x.requires_grad and x.dtype in (torch.float32, torch.bfloat16)


def linear_bf16(x, weight, bias=None, relu=False):
    """``F.relu?(F.linear(x, weight, bias))`` with bf16 operands / output and fp32 accumulation.  ``x``: a StagedInput
    (its bf16 copy is used as it is) or a [N, Fin] tensor (cast and zero-padded to a multiple of 8 columns)."""
    fin = weight.shape[1]
    if isinstance(x, StagedInput):
        xs = x.xs
    else:
        ldx = padded_width(fin) if fin <= 256 else (fin + 7) // 8 * 8
        xs = _stage_rows(x, torch.bfloat16, _lib.ACM_BF16, ldx, _stream())
    return LinearBf16.apply(xs, weight, bias, relu)
