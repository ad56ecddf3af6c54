"""Drop-in replacement of the reference module ``models.layers`` (ACM-Pytorch flavour).

``GraphConvolution`` mirrors the reference class one to one -- constructor signature,
attribute names, parameter names/shapes/creation order (so ``parameters()`` order, the
optimizer state and ``state_dict`` keys are identical), RNG draw order of
``reset_parameters`` and ``__repr__`` -- but its forward is ONE autograd function running
the hand-written sm_100a kernels of libacm_b200 (acm_gnn_b200/functional.py).

Reference: ACM-Pytorch/models/layers.py:14-242 (class), :245-285 (MLP).
The ACM-Geometric flavour (LayerNorm branch live, quirk Q1) is in layers_geometric.py.
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.parameter import Parameter

from .functional import (AcmLayerFunction, LayerConfig, StagedInput, default_dtype, linear_bf16, linear_bf16_eligible,
                         padded_width)
from .operator import AcmOperator, cached_operator

device = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")  # same module-level global as the reference (layers.py:10-11)

# struc_low is [nnodes, out_features] in the reference even when structure_info == 0
# (layers.py:63).  Above this many elements an UNUSED struc_low is allocated empty
# (documented deviation: 10 GB per layer at 10 M x 256); the state_dict key is kept.
_LAZY_STRUC_ELEMS = int(os.environ.get("ACMB200_LAZY_STRUC_ELEMS", str(1 << 27)))


class GraphConvolution(nn.Module):
    _FLAVOUR = "pytorch"

    def __init__(self, in_features, out_features, nnodes, model_type, output_layer=0, variant=False,
                 structure_info=0):
        super().__init__()
        if model_type not in ("mlp", "sgc", "gcn"):
            padded_width(out_features)   # fail at construction, not in the first forward (out_features <= 256)
        self.in_features, self.out_features = in_features, out_features
        self.nnodes = nnodes
        self.output_layer, self.model_type = output_layer, model_type
        self.structure_info, self.variant = structure_info, variant
        self.att_low, self.att_high, self.att_mlp = 0, 0, 0

        def fresh(*shape):
            return Parameter(torch.empty(*shape, dtype=torch.float32, device=device))

        self.weight_low, self.weight_high, self.weight_mlp = (fresh(in_features, out_features) for _ in range(3))
        self.att_vec_low, self.att_vec_high, self.att_vec_mlp = (fresh(out_features, 1) for _ in range(3))
        self.layer_norm_low, self.layer_norm_high, self.layer_norm_mlp = (nn.LayerNorm(out_features) for _ in range(3))
        self.layer_norm_struc_low, self.layer_norm_struc_high = nn.LayerNorm(out_features), nn.LayerNorm(out_features)
        self.att_struc_low = fresh(out_features, 1)
        self._lazy_struc = (not structure_info) and nnodes * out_features > _LAZY_STRUC_ELEMS
        self.struc_low = fresh(0 if self._lazy_struc else nnodes, out_features)
        k = 4 if structure_info else 3
        self.att_vec = fresh(k, k)
        self.reset_parameters()
        # knobs come from the environment so that the reference train.py needs no new flags
        self.acm_dtype = default_dtype()
        self.acm_gemm = os.environ.get("ACMB200_GEMM", "auto")
        self.acm_dist = None  # set by acm_gnn_b200.dist for row-partitioned multi-GPU runs
        self.acm_out_dtype = "fp32"  # "bf16": emit bf16 activations for a following ACM layer (models.GCN opts in)

    def reset_parameters(self):
        # draw order of the reference (layers.py:70-92); bounds are 1/sqrt(size(1))
        stdv = 1.0 / math.sqrt(self.weight_mlp.size(1))
        std_att = 1.0 / math.sqrt(self.att_vec_mlp.size(1))
        std_att_vec = 1.0 / math.sqrt(self.att_vec.size(1))
        for w in (self.weight_low, self.weight_high, self.weight_mlp, self.struc_low):
            w.data.uniform_(-stdv, stdv)
        for a in (self.att_vec_high, self.att_vec_low, self.att_vec_mlp, self.att_struc_low):
            a.data.uniform_(-std_att, std_att)
        self.att_vec.data.uniform_(-std_att_vec, std_att_vec)
        for ln in (self.layer_norm_low, self.layer_norm_high, self.layer_norm_mlp,
                   self.layer_norm_struc_low, self.layer_norm_struc_high):
            ln.reset_parameters()

    # -- checkpoints: a lazily allocated (empty) struc_low stays interchangeable with the reference's ---
    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        key = prefix + "struc_low"
        t = state_dict.get(key)
        if t is not None and self._lazy_struc and t.shape[0] != 0:
            # reference checkpoint ([nnodes, F], unused when structure_info == 0): accept and drop it
            state_dict = dict(state_dict)
            state_dict[key] = t.new_empty(0, t.shape[1])
        elif t is not None and not self._lazy_struc and t.shape[0] == 0 and not self._uses_structure():
            # checkpoint written by a lazy layer: keep the values this (full-size, unused) parameter has
            state_dict = dict(state_dict)
            state_dict[key] = self.struc_low.detach()
        return super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    # -- which branches of the reference forward are live -----------------------------------
    def _ln_live(self):
        if self._FLAVOUR == "pytorch":  # layers.py:96,123 compare against names the CLI never passes
            return self.model_type in ("acmgcn+", "acmgcn++")
        return self.model_type in ("acmgcnp", "acmgcnpp")  # ACM-Geometric/layers.py:59,67

    def _uses_structure(self):
        return bool(self.structure_info) and self.model_type not in ("acmgcn", "acmsnowball")

    def forward(self, input, adj_low, adj_high, adj_low_unnormalized):
        mt = self.model_type
        if mt == "mlp":  # layers.py:156-158 (not an ACM path; plain dense product)
            return torch.mm(input, self.weight_mlp)
        if mt in ("sgc", "gcn"):  # layers.py:159-161
            return torch.mm(adj_low, torch.mm(input, self.weight_low))
        if mt == "acmsgc":
            raise NotImplementedError("acmsgc is broken in the reference (GCN passes no nnodes); out of scope")

        op = adj_low if isinstance(adj_low, AcmOperator) else cached_operator(
            adj_low, adj_high, adj_low_unnormalized if self._uses_structure() else None)
        use_struct = self._uses_structure()
        ln_live = self._ln_live()
        if self.att_vec.shape[0] == 4 and not use_struct:
            # reference: acmgcn + structure_info=1 feeds a 4x4 att_vec to attention3 and crashes
            raise RuntimeError("structure_info=1 is only valid with model_type acmgcnp/acmgcnpp")
        cfg = LayerConfig(variant=bool(self.variant), k_channels=4 if use_struct else 3, ln_live=ln_live,
                          out_scale=1.0 if use_struct else 3.0, dtype=self.acm_dtype, gemm=self.acm_gemm,
                          dist=self.acm_dist, layer_key=id(self),
                          out_dtype=self.acm_out_dtype if self.acm_dtype == "bf16" else "fp32")
        ln_flat = ()
        if ln_live:
            lns = [self.layer_norm_low, self.layer_norm_high, self.layer_norm_mlp] + (
                [self.layer_norm_struc_low] if use_struct else [])
            ln_flat = tuple(t for ln in lns for t in (ln.weight, ln.bias))
        y, att = AcmLayerFunction.apply(
            op, cfg, input, self.weight_low, self.weight_high, self.weight_mlp,
            self.att_vec_low, self.att_vec_high, self.att_vec_mlp, self.att_vec,
            self.struc_low if use_struct else None, self.att_struc_low if use_struct else None, *ln_flat)
        # side-effect attributes of the reference forward (layers.py:167,197,211-216,225)
        self.att_low, self.att_high, self.att_mlp = att[:, 0:1], att[:, 1:2], att[:, 2:3]
        if use_struct:
            self.att_struc_vec_low = att[:, 3:4]
        return y

    def __repr__(self):
        return self.__class__.__name__ + " (" + str(self.in_features) + " -> " + str(self.out_features) + ")"


class MLP(nn.Module):
    """Pass-through equivalent of the reference helper (layers.py:245-285): a stack of
    Linear -> relu -> BatchNorm -> dropout blocks with a final Linear.  Not on the ACM hot
    path; kept so that ``from models.layers import GraphConvolution, MLP`` resolves and
    ``mlpX`` keeps its state_dict keys (``lins.N.*``, ``bns.N.*``)."""

    def __init__(self, in_channels, hidden_channels, out_channels, num_layers, dropout=0.5):
        super().__init__()
        self.lins, self.bns = nn.ModuleList(), nn.ModuleList()
        widths = [in_channels] + [hidden_channels] * (num_layers - 1) + [out_channels]
        for i in range(num_layers):
            self.lins.append(nn.Linear(widths[i], widths[i + 1]))
            if i < num_layers - 1 or num_layers == 1:
                self.bns.append(nn.BatchNorm1d(widths[i + 1]))
        self.dropout = dropout
        self.acm_dtype = default_dtype()

    def reset_parameters(self):
        for m in list(self.lins) + list(self.bns):
            m.reset_parameters()

    def forward(self, data, input_tensor=False, relu_out=False):
        """``relu_out`` (ours): fold the caller's ``F.relu`` of the result into the last Linear.

        In bf16 storage mode (``ACMB200_DTYPE=bf16``) a Linear whose input needs no gradient -- the
        single-Linear ``mlpX`` of acmgcn++ on the raw features -- runs on the tcgen05 GEMM with bf16
        operands and an fp32 bias/relu epilogue (functional.linear_bf16) and returns bf16; in fp32 mode
        (the default) it is torch's fp32 ``F.linear``, the reference's arithmetic."""
        x = data if input_tensor else data.graph["node_feat"]
        for lin, bn in zip(self.lins[:-1], self.bns):
            x = F.dropout(bn(F.relu(self._lin(lin, x), inplace=True)), p=self.dropout, training=self.training)
        return self._lin(self.lins[-1], x, relu_out)

    def _lin(self, lin, x, relu=False):
        if self.acm_dtype == "bf16" and os.environ.get("ACMB200_LINEAR", "auto") != "off" and linear_bf16_eligible(x, lin.weight):
            return linear_bf16(x, lin.weight, lin.bias, relu)
        if isinstance(x, StagedInput):
            x = x.x
        y = lin(x)
        return F.relu(y) if relu else y
