"""Whole train step / eval forward captured ONCE in a CUDA graph and replayed.

On the graphs the reference actually ships (Cora: 2 708 nodes, Squirrel: 5 201 nodes; BASELINE
configs 1-2) one train step of the 2-layer model is ~20 launches of libacm_b200 plus torch's glue
and optimizer kernels, each a few microseconds of GPU work: the step is bound by host launch
overhead (~1.6 ms measured), not by the kernels.  Every launch of the library goes to the
current torch stream, never synchronises and never allocates (include/acm_b200.h), so the whole
step -- ``zero_grad``, forward, fused log-softmax/NLL, backward, ``optimizer.step()`` -- can be
captured in one ``cudaGraph`` and replayed with a single launch.

This is the train step of the reference drivers (ACM-Pytorch/utils.py:547-574 ``train_model``;
ACM-Geometric/train.py:120-136) and their eval forward (ACM-Pytorch/train.py:110-121) for a
FIXED set of inputs: full-batch training passes the same features, operator, labels and split
mask every epoch.  New data (another split, new features) is written into the static tensors
in place (``step.x.copy_(...)`` -- ``step.x.update_(...)`` for a StagedInput --, ``step.labels.copy_(...)``,
``step.train_mask.copy_(...)``).

Requirements: single GPU (no row partition attached); an optimizer constructed with
``capturable=True`` (Adam / AdamW, the reference's optimizers); dropout is supported (torch's
Philox generator is graph-safe) but its mask sequence differs from an eager run of the same
seed.  Warm-up steps run eagerly before the capture and are rolled back, so the captured step
starts from exactly the parameters and optimizer state it was handed.
"""
from __future__ import annotations

import torch

from . import _lib
from .functional import StagedInput, nll_log_softmax
from .layers import GraphConvolution
from .operator import AcmOperator, cached_operator


def _check_single_gpu(model):
    for m in model.modules():
        if isinstance(m, GraphConvolution) and m.acm_dist is not None:
            raise NotImplementedError("CUDA-graph capture of a row-partitioned model is not supported "
                                      "(the exchange barriers are host-driven); run it eagerly")


def _tensors_of(x):
    return x.x if isinstance(x, StagedInput) else x


def _resolve_adj(model, adj):
    """The captured graph bakes the device pointers of the operator's CSR arrays in, so the operator is
    resolved ONCE here and held by a strong reference for the life of the graph (the conversion cache
    of operator.cached_operator may evict its own reference at any time)."""
    adj = tuple(adj)
    if isinstance(adj[0], AcmOperator):
        return (adj[0], None, None)
    uses_struct = any(isinstance(m, GraphConvolution) and m._uses_structure() for m in model.modules())
    return (cached_operator(adj[0], adj[1], adj[2] if uses_struct else None), None, None)


class _NoTimer:
    """The per-launch CUDA-event timer of bench.py cannot record inside a capture."""

    def __enter__(self):
        self.saved = _lib._TIMER
        _lib.set_timer(None)

    def __exit__(self, *exc):
        _lib.set_timer(self.saved)


class GraphedTrainStep:
    """``loss = step()`` replays zero_grad + forward + masked mean NLL + backward + optimizer.step.

    model      : acm_gnn_b200.GCN (or the reference's GCN running on the drop-in layer)
    optimizer  : built with capturable=True over the model's parameters
    x          : [N, Fin] fp32 CUDA tensor (static; new values with ``step.x.copy_(...)``) or a StagedInput
                 (new values ONLY through ``step.x.update_(new_x)``, which refreshes the kernel-layout
                 copies the captured kernels read)
    adj        : (adj_low, adj_high, adj_low_unnormalized) exactly as the reference passes them,
                 or (AcmOperator, None, None)
    labels     : int64 [N];  train_mask : uint8/bool [N] (1 = training row), the dense form of
                 ``idx_train`` (utils.py:567-568)
    """

    def __init__(self, model, optimizer, x, adj, labels, train_mask, warmup: int = 3):
        if not _tensors_of(x).is_cuda:
            raise RuntimeError("acm_gnn_b200: CUDA tensors only (there is no CPU fallback)")
        _check_single_gpu(model)
        for grp in optimizer.param_groups:
            if "capturable" in grp and not grp["capturable"]:
                raise ValueError("GraphedTrainStep needs an optimizer built with capturable=True "
                                 "(torch.optim.Adam/AdamW(..., capturable=True))")
        self.model, self.optimizer = model, optimizer
        self.x, self.adj = x, _resolve_adj(model, adj)
        self.labels = labels.to(torch.int64).contiguous()
        self.train_mask = train_mask.to(torch.uint8).contiguous()
        # the normaliser 1/|train| is a launch argument (host scalar) and therefore frozen in the
        # graph: a new mask with a different count needs a new GraphedTrainStep
        self.n_train = int(self.train_mask.sum().item())
        self.launches_per_step = 0
        self.loss = None
        self._capture(max(int(warmup), 1))

    # one eager step, exactly what gets captured
    def _step(self):
        self.optimizer.zero_grad(set_to_none=True)
        out = self.model(self.x, *self.adj)
        loss = nll_log_softmax(out, self.labels, self.train_mask, n_train=self.n_train)
        loss.backward()
        self.optimizer.step()
        return out, loss

    def _snapshot(self):
        params = [p.detach().clone() for p in self.model.parameters()]
        bufs = [b.detach().clone() for b in self.model.buffers()]
        state = {}
        for p, st in self.optimizer.state.items():
            state[p] = {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()}
        return params, bufs, state

    def _restore(self, snap):
        params, bufs, state = snap
        with torch.no_grad():
            for p, s in zip(self.model.parameters(), params):
                p.copy_(s)
            for b, s in zip(self.model.buffers(), bufs):
                b.copy_(s)
            for p, st in self.optimizer.state.items():
                old = state.get(p)
                for k, v in st.items():
                    if not torch.is_tensor(v):
                        continue
                    if old is not None and k in old:
                        v.copy_(old[k])
                    else:
                        v.zero_()      # state created by the warm-up: Adam/AdamW start from zeros

    def _capture(self, warmup):
        self.model.train()
        snap = self._snapshot()
        rng = torch.cuda.get_rng_state()
        with _NoTimer():
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):      # builds the operator cache, the optimizer state, cuBLAS handles ...
                    self._step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self._restore(snap)
            torch.cuda.set_rng_state(rng)
            self.optimizer.zero_grad(set_to_none=True)
            self.graph = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(self.graph):
                self.out, self.loss = self._step()
            self.launches_per_step = _lib.launch_count() - n0
        self.loss = self.loss.detach()
        self.out = self.out.detach()

    def __call__(self):
        """Replay one train step; returns the (static) loss tensor of that step."""
        self.graph.replay()
        return self.loss


class GraphedForward:
    """Eval-mode forward (``model.eval(); output = model(...)``, ACM-Pytorch/train.py:110-111)
    captured under ``torch.no_grad()``; ``out = fwd()`` replays it and returns the static output."""

    def __init__(self, model, x, adj, warmup: int = 2):
        if not _tensors_of(x).is_cuda:
            raise RuntimeError("acm_gnn_b200: CUDA tensors only (there is no CPU fallback)")
        _check_single_gpu(model)
        self.model, self.x, self.adj = model, x, _resolve_adj(model, adj)
        was_training = model.training
        model.eval()
        with _NoTimer(), torch.no_grad():
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(max(int(warmup), 1)):
                    model(self.x, *self.adj)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(self.graph):
                self.out = model(self.x, *self.adj)
            self.launches_per_step = _lib.launch_count() - n0
        model.train(was_training)

    def __call__(self):
        self.graph.replay()
        return self.out
