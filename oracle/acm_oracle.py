"""CPU oracle for the ACM graph-convolution hot path  --  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, the algorithm of the reference hot path
(SitaoLuan/ACM-GNN).  It is the *checker* for the CUDA product path in
``acm_gnn_b200``; nothing in the product imports it.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py`` (its ``cpu_baseline`` leg and
``--impl reference``) may import this module.

Parity status: **pinned**.  The reference has no tests or golden vectors of its own
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference itself,
run in the build container: ``tests/golden/make_golden.py`` imports the unmodified
reference modules from ``/root/reference`` and stores their forward outputs,
attention columns and all gradients under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this restatement against those files.

Arithmetic follows the reference: fp32 everywhere, ``torch.mm`` for dense products and
``torch.sparse.mm`` on COO operands for the aggregations (the "CPU torch.sparse.mm path"
BASELINE.json names), gradients by autograd.  Integer work (CSR indices, degrees) is numpy.

Reference citations are relative to ``/root/reference``:
  * operator construction  ACM-Pytorch/utils.py:421-438, 619-629   (fp32, dense)
                           ACM-Geometric/utils.py:5-28, train.py:66-84 (fp64 scipy, cast)
  * layer forward          ACM-Pytorch/models/layers.py:154-232 ; attention 94-152
                           ACM-Geometric/layers.py:57-116 (LayerNorm branch is live here)
  * layer stack            ACM-Pytorch/models/models.py:106-166
  * train step             ACM-Pytorch/utils.py:547-574
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# Operator construction (integer / degree work: numpy, bit-exact targets)
# --------------------------------------------------------------------------------------


@dataclass
class CsrOperator:
    """``A_low = D^-1 (A + I)`` in CSR with strictly increasing columns per row."""

    n: int
    rowptr: np.ndarray  # int64 [n+1]
    col: np.ndarray  # int64 [nnz]
    mult: np.ndarray  # float32 [nnz]  entries m_ij of (A + I), duplicates summed
    rowsum: np.ndarray  # float32 [n]
    rinv: np.ndarray  # float32 [n]   1/rowsum, inf -> 0
    w_low: np.ndarray  # float32 [nnz]  fl32(rinv_i * m_ij)      (the adj_low values)
    w_high: np.ndarray  # float32 [nnz]  [i==j] - w_low           (the adj_high values)

    @property
    def nnz(self) -> int:
        return int(self.col.shape[0])

    def rows(self) -> np.ndarray:
        return np.repeat(np.arange(self.n, dtype=np.int64), np.diff(self.rowptr))


def symmetrise_edges(row: np.ndarray, col: np.ndarray, n: int) -> Tuple[np.ndarray, np.ndarray]:
    """PyG ``to_undirected`` as used at ACM-Geometric/train.py:66-67: concatenate the edge
    list with its reverse, sort by (row, col) and drop duplicates (self-loops are kept)."""
    r = np.concatenate([row, col]).astype(np.int64)
    c = np.concatenate([col, row]).astype(np.int64)
    key = np.unique(r * n + c)
    return key // n, key % n


def _coalesce_plus_identity(row, col, n, val=None):
    """CSR pattern and summed multiplicities of ``I + A`` (duplicates in A are summed, as
    ``to_dense()`` / scipy's COO->CSR conversion do)."""
    row = np.asarray(row, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    v = np.ones(row.shape[0], dtype=np.float64) if val is None else np.asarray(val, np.float64)
    r = np.concatenate([row, np.arange(n, dtype=np.int64)])
    c = np.concatenate([col, np.arange(n, dtype=np.int64)])
    v = np.concatenate([v, np.ones(n, dtype=np.float64)])
    key = r * n + c
    ukey, inv = np.unique(key, return_inverse=True)
    mult = np.zeros(ukey.shape[0], dtype=np.float64)
    np.add.at(mult, inv, v)
    urow = ukey // n
    ucol = ukey % n
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr, urow + 1, 1)
    rowptr = np.cumsum(rowptr)
    return rowptr, ucol, urow, mult


def build_operator(row, col, n: int, flavour: str = "pytorch", val=None) -> CsrOperator:
    """Row-normalised low-pass operator and its high-pass complement.

    flavour "pytorch":   ACM-Pytorch/utils.py:421-438,626-628 -- everything in fp32:
        rowsum = sum_j m_ij ; rinv = rowsum**-1 (inf->0) ; w = fl32(rinv_i*m_ij) ;
        high = fl32([i==j] - w).
    flavour "geometric": ACM-Geometric/utils.py:5-19 + train.py:76-80 -- the same formula in
        fp64 (scipy), ``I - A_low`` in fp64, both cast to fp32 afterwards (utils.py:23).
    """
    rowptr, ucol, urow, mult64 = _coalesce_plus_identity(row, col, n, val)
    diag = (urow == ucol)
    if flavour == "pytorch":
        mult = mult64.astype(np.float32)
        rowsum = np.zeros(n, dtype=np.float32)
        # fp32 row sums of small integers are exact irrespective of order
        np.add.at(rowsum, urow, mult)
        with np.errstate(divide="ignore"):
            rinv = (np.float32(1.0) / rowsum).astype(np.float32)
        rinv[np.isinf(rinv)] = 0.0
        w_low = (rinv[urow] * mult).astype(np.float32)
        w_high = (diag.astype(np.float32) - w_low).astype(np.float32)
    elif flavour == "geometric":
        rowsum64 = np.zeros(n, dtype=np.float64)
        np.add.at(rowsum64, urow, mult64)
        with np.errstate(divide="ignore"):
            rinv64 = np.power(rowsum64, -1.0)
        rinv64[np.isinf(rinv64)] = 0.0
        w64 = rinv64[urow] * mult64
        w_low = w64.astype(np.float32)
        w_high = (diag.astype(np.float64) - w64).astype(np.float32)
        rowsum = rowsum64.astype(np.float32)
        rinv = rinv64.astype(np.float32)
        mult = mult64.astype(np.float32)
    else:
        raise ValueError(flavour)
    return CsrOperator(n, rowptr, ucol, mult, rowsum, rinv, w_low, w_high)


def operator_to_torch(op: CsrOperator, dense_low: bool = False):
    """The tensors the reference drivers hand to ``GCN.forward``: ``adj_low`` (dense fp32 in
    the ACM-Pytorch flavour, sparse COO in the Geometric one) and ``adj_high`` (sparse COO
    with exact zeros dropped, as ``.to_sparse()`` / scipy do)."""
    rows = torch.from_numpy(op.rows())
    cols = torch.from_numpy(op.col.astype(np.int64))
    idx = torch.stack([rows, cols])
    low = torch.sparse_coo_tensor(idx, torch.from_numpy(op.w_low), (op.n, op.n)).coalesce()
    keep = torch.from_numpy(op.w_high != 0)
    high = torch.sparse_coo_tensor(idx[:, keep], torch.from_numpy(op.w_high)[keep], (op.n, op.n)).coalesce()
    if dense_low:
        low = low.to_dense()
    return low, high


def raw_adjacency_to_torch(row, col, n: int):
    """``adj_low_unnormalized``: the raw adjacency as sparse COO of ones (duplicates summed
    when consumed), ACM-Pytorch/utils.py:611,624 ; ACM-Geometric/train.py:76,81."""
    idx = torch.from_numpy(np.stack([np.asarray(row, np.int64), np.asarray(col, np.int64)]))
    return torch.sparse_coo_tensor(idx, torch.ones(idx.shape[1], dtype=torch.float32), (n, n)).coalesce()


def row_normalise_features(x: torch.Tensor) -> torch.Tensor:
    """Feature L1 row normalisation, ACM-Pytorch/utils.py:612-617 (normalize_tensor)."""
    rowsum = x.sum(1)
    rinv = rowsum.pow(-1)
    rinv[torch.isinf(rinv)] = 0.0
    return rinv[:, None] * x


# --------------------------------------------------------------------------------------
# Parameters (names and draw order are the reference's state_dict contract)
# --------------------------------------------------------------------------------------

LN_NAMES = ("layer_norm_low", "layer_norm_high", "layer_norm_mlp", "layer_norm_struc_low", "layer_norm_struc_high")


def init_layer_params(in_features: int, out_features: int, nnodes: int, structure_info: int = 0,
                      generator: Optional[torch.Generator] = None) -> Dict[str, torch.Tensor]:
    """Same-seed initialisation in the reference's RNG draw order
    (ACM-Pytorch/models/layers.py:70-92): weight_low, weight_high, weight_mlp, struc_low,
    att_vec_high, att_vec_low, att_vec_mlp, att_struc_low, att_vec."""
    k = 4 if structure_info else 3
    stdv = 1.0 / math.sqrt(out_features)
    std_att = 1.0  # 1/sqrt(att_vec_mlp.size(1)) with size(1) == 1
    std_att_vec = 1.0 / math.sqrt(k)
    p: Dict[str, torch.Tensor] = {}

    def u(shape, a):
        return torch.empty(shape, dtype=torch.float32).uniform_(-a, a, generator=generator)

    p["weight_low"] = u((in_features, out_features), stdv)
    p["weight_high"] = u((in_features, out_features), stdv)
    p["weight_mlp"] = u((in_features, out_features), stdv)
    p["struc_low"] = u((nnodes, out_features), stdv)
    p["att_vec_high"] = u((out_features, 1), std_att)
    p["att_vec_low"] = u((out_features, 1), std_att)
    p["att_vec_mlp"] = u((out_features, 1), std_att)
    p["att_struc_low"] = u((out_features, 1), std_att)
    p["att_vec"] = u((k, k), std_att_vec)
    for name in LN_NAMES:
        p[name + ".weight"] = torch.ones(out_features)
        p[name + ".bias"] = torch.zeros(out_features)
    return p


# --------------------------------------------------------------------------------------
# Layer forward (float math: torch CPU, fp32)
# --------------------------------------------------------------------------------------


def _spmm(adj, dense):
    if adj.layout == torch.strided:
        return torch.mm(adj, dense)
    return torch.sparse.mm(adj, dense)


def layer_norm_is_live(model_type: str, flavour: str) -> bool:
    """Quirk Q1.  ACM-Pytorch tests the strings "acmgcn+"/"acmgcn++" (layers.py:96,123) which
    its CLI never produces -> LayerNorm only runs if a caller passes those literal names;
    ACM-Geometric tests "acmgcnp"/"acmgcnpp" (layers.py:59,67) -> live."""
    if flavour == "pytorch":
        return model_type in ("acmgcn+", "acmgcn++")
    return model_type in ("acmgcnp", "acmgcnpp")


def layer_forward(p: Dict[str, torch.Tensor], x: torch.Tensor, adj_low, adj_high, adj_low_unnormalized,
                  model_type: str = "acmgcn", variant=False, structure_info: int = 0,
                  flavour: str = "pytorch"):
    """One ACM layer.  Returns ``(Y, att)`` with ``att`` the ``[N,3|4]`` channel weights.

    Follows ACM-Pytorch/models/layers.py:176-232: three feature transforms, low-pass and
    high-pass aggregation (relu before aggregation when ``variant`` else after), identity
    channel, optional structure channel, sigmoid/softmax channel attention and the mix."""
    hl = torch.mm(x, p["weight_low"])
    hh = torch.mm(x, p["weight_high"])
    hi = torch.mm(x, p["weight_mlp"])
    if variant:
        o_l = _spmm(adj_low, F.relu(hl))
        o_h = _spmm(adj_high, F.relu(hh))
    else:
        o_l = F.relu(_spmm(adj_low, hl))
        o_h = F.relu(_spmm(adj_high, hh))
    o_i = F.relu(hi)
    chans = [o_l, o_h, o_i]
    avs = [p["att_vec_low"], p["att_vec_high"], p["att_vec_mlp"]]
    lns = ["layer_norm_low", "layer_norm_high", "layer_norm_mlp"]
    use_struct = bool(structure_info) and model_type not in ("acmgcn", "acmsnowball")
    if use_struct:
        o_s = F.relu(_spmm(adj_low_unnormalized, p["struc_low"]))
        chans.append(o_s)
        avs.append(p["att_struc_low"])
        lns.append("layer_norm_struc_low")
    k = len(chans)
    live = layer_norm_is_live(model_type, flavour)
    zs = []
    for o, a, ln in zip(chans, avs, lns):
        if live:
            o = F.layer_norm(o, (o.shape[1],), p[ln + ".weight"], p[ln + ".bias"], 1e-5)
        zs.append(torch.mm(o, a))
    s = torch.sigmoid(torch.cat(zs, 1))
    logits = torch.mm(s, p["att_vec"]) / k
    att = torch.softmax(logits, 1)
    scale = 1.0 if use_struct else 3.0
    y = chans[0] * att[:, 0:1]
    for j in range(1, k):
        y = y + chans[j] * att[:, j:j + 1]
    return scale * y, att


def gcn_forward(params: Dict[str, Dict[str, torch.Tensor]], x, adj_low, adj_high, adj_low_unnormalized,
                model_type="acmgcn", variant=False, structure_info=0, flavour="pytorch",
                dropout: float = 0.0, training: bool = True):
    """Layer stack, ACM-Pytorch/models/models.py:106-166 (acmgcn / acmgcnp / acmgcnpp):
    dropout -> [mlpX residual branch] -> layer 0 -> relu -> dropout -> layer 1."""
    x = F.dropout(x, dropout, training=training)
    xx = None
    if model_type == "acmgcnpp":
        lin = params["mlpX"]
        xx = F.dropout(F.relu(F.linear(x, lin["weight"], lin["bias"])), dropout, training=training)
    kw = dict(model_type=model_type, variant=variant, structure_info=structure_info, flavour=flavour)
    f1, att0 = layer_forward(params["gcns.0"], x, adj_low, adj_high, adj_low_unnormalized, **kw)
    f1 = F.dropout(F.relu(f1), dropout, training=training)
    if xx is not None:
        f1 = f1 + xx
    f2, att1 = layer_forward(params["gcns.1"], f1, adj_low, adj_high, adj_low_unnormalized, **kw)
    return f2, (att0, att1)


def train_step_loss(out: torch.Tensor, labels: torch.Tensor, idx_train: torch.Tensor) -> torch.Tensor:
    """log_softmax + NLL on the training rows, ACM-Pytorch/utils.py:567-568."""
    return F.nll_loss(F.log_softmax(out, dim=1)[idx_train], labels[idx_train])


def init_gcn_params(nfeat, nhid, nclass, nnodes, model_type="acmgcn", structure_info=0, generator=None):
    """Replicated-parameter init in the order ``GCN.__init__`` creates modules
    (ACM-Pytorch/models/models.py:39-75): optional mlpX Linear, then layer 0, layer 1."""
    params: Dict[str, Dict[str, torch.Tensor]] = {}
    if model_type == "acmgcnpp":
        bound = 1.0 / math.sqrt(nfeat)
        w = torch.empty(nhid, nfeat)
        torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5), generator=generator)
        b = torch.empty(nhid).uniform_(-bound, bound, generator=generator)
        params["mlpX"] = {"weight": w, "bias": b}
    params["gcns.0"] = init_layer_params(nfeat, nhid, nnodes, structure_info, generator)
    params["gcns.1"] = init_layer_params(nhid, nclass, nnodes, structure_info, generator)
    return params


# --------------------------------------------------------------------------------------
# Synthetic graphs
# --------------------------------------------------------------------------------------


def synthetic_edges(n: int, e_directed: int, seed: int = 0, zipf: float = 0.0):
    """Uniform random undirected pairs -> symmetric directed edge list without self loops.
    ``zipf`` > 0 skews the destination distribution (degree-imbalance test graphs)."""
    rng = np.random.default_rng(seed)
    eu = e_directed // 2
    src = rng.integers(0, n, eu, dtype=np.int64)
    if zipf > 0:
        dst = np.minimum((rng.zipf(1.0 + zipf, eu) - 1), n - 1).astype(np.int64)
        perm = rng.permutation(n)
        dst = perm[dst]
    else:
        dst = rng.integers(0, n, eu, dtype=np.int64)
    keep = src != dst
    return symmetrise_edges(src[keep], dst[keep], n)
