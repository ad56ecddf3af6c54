"""Device-resident CSR form of the ACM aggregation operators.

The reference hands ``GCN.forward`` the low-pass operator ``A_low = D^-1 (A + I)`` either
as a DENSE fp32 ``[N,N]`` tensor (ACM-Pytorch/utils.py:626-628) or as sparse COO
(ACM-Geometric/train.py:77-79), plus a separately materialised ``A_high = I - A_low``
(sparse COO).  The kernels consume ONE CSR per operator: ``rowptr`` int64, ``col`` int32,
``val`` fp32 (exactly the reference's ``adj_low`` values) and the CSR of the transpose for
the backward pass.  ``A_high`` is never materialised: ``A_high Z = Z - A_low Z``; the tensor
the driver passes is only validated once against ``I - A_low``.
"""
from __future__ import annotations

import collections
import weakref
from typing import Optional

import torch

from . import _lib

_I32_MAX = 2**31 - 1


def _stream():
    return torch.cuda.current_stream().cuda_stream


class CsrMatrix:
    """CSR arrays on one device + the CSR of the transpose (shared index arrays when the
    sparsity pattern is symmetric)."""

    def __init__(self, n_rows, n_cols, rowptr, col, val):
        self.n_rows, self.n_cols = int(n_rows), int(n_cols)
        self.rowptr, self.col, self.val = rowptr, col, val
        self.rowptr_t = self.col_t = self.val_t = None
        self.symmetric_pattern = None

    @property
    def nnz(self):
        return int(self.col.shape[0])

    @property
    def device(self):
        return self.col.device

    # -- construction -------------------------------------------------------------------
    @staticmethod
    def from_sorted_coo(row, col, val, n_rows, n_cols):
        """``row`` int64 ascending (row-major sorted, duplicates already summed)."""
        dev = row.device
        nnz = int(row.shape[0])
        if n_cols > _I32_MAX:
            raise ValueError("column ids must fit int32")
        rowptr = torch.empty(n_rows + 1, dtype=torch.int64, device=dev)
        _lib.call("acm_csr_rowptr", _lib.ptr(row), nnz, n_rows, _lib.ptr(rowptr), _stream())
        return CsrMatrix(n_rows, n_cols, rowptr, col.to(torch.int32), val.to(torch.float32).contiguous())

    def build_transpose(self):
        """Fill ``rowptr_t/col_t/val_t``.  Symmetric pattern (undirected graphs, the
        reference default): same index arrays, values permuted on the GPU by
        ``acm_csr_transpose_values``.  Otherwise (``--directed``): explicit re-sort."""
        if self.rowptr_t is not None:
            return self
        dev = self.device
        if self.n_rows == self.n_cols:
            val_t = torch.empty_like(self.val)
            flag = torch.zeros(1, dtype=torch.int32, device=dev)
            _lib.call("acm_csr_transpose_values", _lib.ptr(self.rowptr), _lib.ptr(self.col), _lib.ptr(self.val),
                      self.n_rows, _lib.ptr(val_t), _lib.ptr(flag), _stream())
            if int(flag.item()) == 0:
                self.rowptr_t, self.col_t, self.val_t = self.rowptr, self.col, val_t
                self.symmetric_pattern = True
                return self
        self.symmetric_pattern = False
        rows = torch.repeat_interleave(torch.arange(self.n_rows, device=dev, dtype=torch.int64),
                                       self.rowptr[1:] - self.rowptr[:-1])
        key = self.col.to(torch.int64) * self.n_rows + rows
        order = torch.argsort(key)
        t = CsrMatrix.from_sorted_coo(self.col.to(torch.int64)[order], rows[order], self.val[order],
                                      self.n_cols, self.n_rows)
        self.rowptr_t, self.col_t, self.val_t = t.rowptr, t.col, t.val
        return self

    # -- degree skew: segment tables of the long rows (see csrc/acm_common.cuh kLongRow) ---------
    LONG_ROW = 256

    def long_rows(self, transposed=False):
        """(rows int32 [L] ascending, seg_long int32 [S], seg_e0 int64 [S], seg_e1 int64 [S]) of
        the rows with more than LONG_ROW stored edges, or None when there are none."""
        attr = "_long_t" if transposed else "_long"
        if hasattr(self, attr):
            return getattr(self, attr)
        rowptr = self.rowptr_t if transposed else self.rowptr
        deg = rowptr[1:] - rowptr[:-1]
        long = torch.nonzero(deg > self.LONG_ROW).squeeze(1)
        res = None
        if long.numel() > 0:
            t = self.LONG_ROW
            nseg = (deg[long] + t - 1) // t
            seg_long = torch.repeat_interleave(torch.arange(long.numel(), device=long.device), nseg)
            first = torch.cumsum(nseg, 0) - nseg
            within = torch.arange(seg_long.numel(), device=long.device) - first[seg_long]
            e0 = rowptr[long][seg_long] + within * t
            e1 = torch.minimum(e0 + t, rowptr[long + 1][seg_long])
            res = (long.to(torch.int32).contiguous(), seg_long.to(torch.int32).contiguous(),
                   e0.contiguous(), e1.contiguous())
        setattr(self, attr, res)
        return res

    # -- narrow rows: processing order that groups rows of equal degree ----------------------------
    ORDER_WINDOW = 4096

    def row_order(self, transposed=False):
        """int32 permutation of the local rows: inside every window of ORDER_WINDOW consecutive rows
        the rows are sorted by their number of stored edges (stable).  The gather kernels put
        32/(fp/8) rows in one warp and walk max(degree) edges, so neighbours of equal degree remove
        the divergence (on the 10 M-node uniform graph, 16 rows per warp, 65 % of the lanes are
        active in the gather loop), while the windows keep the streamed per-row accesses local.
        OPT-IN (``ACMB200_ROW_ORDER=1``): measured on that graph it does not pay -- 4.63 vs 4.55 ms
        forward, 4.88 vs 4.87 ms backward -- because those kernels are bound by the rate of random
        64-byte DRAM accesses (4.1 TB/s), not by lane utilisation; kept for strongly skewed graphs.
        Built once per operator."""
        attr = "_order_t" if transposed else "_order"
        if hasattr(self, attr):
            return getattr(self, attr)
        import os
        res = None
        if os.environ.get("ACMB200_ROW_ORDER", "0") == "1" and self.n_rows > 1:
            rowptr = self.rowptr_t if transposed else self.rowptr
            deg = rowptr[1:] - rowptr[:-1]
            win = torch.arange(self.n_rows, device=deg.device, dtype=torch.int64) // self.ORDER_WINDOW
            key = win * (int(deg.max().item()) + 1) + deg
            res = torch.argsort(key, stable=True).to(torch.int32).contiguous()
        setattr(self, attr, res)
        return res

    def rows(self):
        return torch.repeat_interleave(torch.arange(self.n_rows, device=self.device, dtype=torch.int64),
                                       self.rowptr[1:] - self.rowptr[:-1])

    def row_slice(self, r0, r1):
        """Rows [r0, r1) with GLOBAL column ids (1-D row partition of the operator)."""
        e0, e1 = int(self.rowptr[r0].item()), int(self.rowptr[r1].item())
        m = CsrMatrix(r1 - r0, self.n_cols, (self.rowptr[r0:r1 + 1] - e0).contiguous(),
                      self.col[e0:e1].contiguous(), self.val[e0:e1].contiguous())
        if self.rowptr_t is not None:
            t0, t1 = int(self.rowptr_t[r0].item()), int(self.rowptr_t[r1].item())
            m.rowptr_t = (self.rowptr_t[r0:r1 + 1] - t0).contiguous()
            m.col_t = self.col_t[t0:t1].contiguous()
            m.val_t = self.val_t[t0:t1].contiguous()
            m.symmetric_pattern = self.symmetric_pattern
        return m


def _coo_parts(adj):
    """(row, col, val) int64/int64/fp32, row-major sorted and coalesced, of a dense or
    sparse-COO torch tensor."""
    if adj.layout == torch.strided:
        adj = adj.to_sparse()
    elif adj.layout != torch.sparse_coo:
        adj = adj.to_sparse_coo()
    adj = adj.coalesce()
    idx = adj.indices()
    return idx[0].contiguous(), idx[1].contiguous(), adj.values().to(torch.float32)


class AcmOperator:
    """``A_low`` (CSR + transpose) and, optionally, the raw adjacency of the structure channel."""

    def __init__(self, low: CsrMatrix, raw: Optional[CsrMatrix] = None):
        self.low = low
        self.raw = raw
        self.n = low.n_rows
        self.row0 = 0           # global id of the first local row (1-D row partition)
        self.n_global = low.n_cols

    @property
    def nnz(self):
        return self.low.nnz

    # -- from what the reference driver passes --------------------------------------------
    @staticmethod
    def from_adjacency(adj_low, adj_high=None, adj_low_unnormalized=None, validate=True):
        if not adj_low.is_cuda:
            raise RuntimeError("acm_gnn_b200 runs on CUDA tensors only (no CPU fallback)")
        n = adj_low.shape[0]
        row, col, val = _coo_parts(adj_low)
        low = CsrMatrix.from_sorted_coo(row, col, val, n, adj_low.shape[1]).build_transpose()
        if validate and adj_high is not None:
            _validate_high(row, col, val, adj_high, n)
        raw = None
        if adj_low_unnormalized is not None:
            r, c, v = _coo_parts(adj_low_unnormalized)
            raw = CsrMatrix.from_sorted_coo(r, c, v, n, adj_low_unnormalized.shape[1]).build_transpose()
        return AcmOperator(low, raw)

    # -- from an edge list (large graphs: no dense [N,N], no O(N^3) diag product) ---------
    @staticmethod
    def from_edges(row, col, n, flavour="pytorch", with_raw=False, edge_val=None):
        """``D^-1 (A + I)`` from the directed edge list of A (duplicates are summed, as
        ``to_dense()`` / scipy do).  Bit-exact with ACM-Pytorch/utils.py:421-438,626-628
        (flavour "pytorch": fp32 reciprocal and product) or ACM-Geometric/utils.py:5-19
        (flavour "geometric": fp64, then cast)."""
        dev = row.device
        if dev.type != "cuda":
            raise RuntimeError("acm_gnn_b200 runs on CUDA tensors only (no CPU fallback)")
        row = row.to(torch.int64)
        col = col.to(torch.int64)
        ar = torch.arange(n, device=dev, dtype=torch.int64)
        key = torch.cat([row * n + col, ar * n + ar])
        if edge_val is None:
            ukey, cnt = torch.unique(key, return_counts=True)  # sorted
            mult64 = cnt.to(torch.float64)
        else:
            ukey, inv = torch.unique(key, return_inverse=True)
            mult64 = torch.zeros(ukey.shape[0], dtype=torch.float64, device=dev)
            mult64.index_add_(0, inv, torch.cat([edge_val.to(torch.float64), torch.ones(n, dtype=torch.float64, device=dev)]))
        del key
        urow = torch.div(ukey, n, rounding_mode="floor")
        ucol = ukey - urow * n
        del ukey
        low = CsrMatrix.from_sorted_coo(urow, ucol, mult64.to(torch.float32), n, n)
        if flavour == "pytorch":
            w = torch.empty_like(low.val)
            rowsum = torch.empty(n, dtype=torch.float32, device=dev)
            rinv = torch.empty(n, dtype=torch.float32, device=dev)
            _lib.call("acm_degree_normalise", _lib.ptr(low.rowptr), _lib.ptr(low.val), n,
                      _lib.ptr(rowsum), _lib.ptr(rinv), _lib.ptr(w), _stream())
            low.val = w
        elif flavour == "geometric":
            rowsum64 = torch.zeros(n, dtype=torch.float64, device=dev).index_add_(0, urow, mult64)
            rinv64 = rowsum64.pow(-1.0)
            rinv64[torch.isinf(rinv64)] = 0.0
            low.val = (rinv64[urow] * mult64).to(torch.float32)
            rinv = rinv64.to(torch.float32)
            rowsum = rowsum64.to(torch.float32)
        else:
            raise ValueError(flavour)
        del urow, ucol, mult64
        low.build_transpose()
        raw = None
        if with_raw:
            k2 = row * n + col
            if edge_val is None:
                uk, cnt = torch.unique(k2, return_counts=True)
                v = cnt.to(torch.float32)
            else:
                uk, inv = torch.unique(k2, return_inverse=True)
                v = torch.zeros(uk.shape[0], dtype=torch.float32, device=dev).index_add_(0, inv, edge_val.to(torch.float32))
            r = torch.div(uk, n, rounding_mode="floor")
            raw = CsrMatrix.from_sorted_coo(r, uk - r * n, v, n, n).build_transpose()
        op = AcmOperator(low, raw)
        op.rinv, op.rowsum = rinv, rowsum
        return op

    def partition(self, r0, r1):
        """Local view for a 1-D row partition: rows [r0,r1), global column ids."""
        p = AcmOperator(self.low.row_slice(r0, r1), self.raw.row_slice(r0, r1) if self.raw is not None else None)
        p.row0, p.n_global = r0, self.n_global
        return p

    # -- torch views (tests, interop) -------------------------------------------------------
    def to_torch_coo(self):
        idx = torch.stack([self.low.rows(), self.low.col.to(torch.int64)])
        return torch.sparse_coo_tensor(idx, self.low.val, (self.low.n_rows, self.low.n_cols)).coalesce()

    def high_to_torch_coo(self):
        """``I - A_low`` as the reference materialises it (exact zeros dropped)."""
        rows, cols = self.low.rows() + self.row0, self.low.col.to(torch.int64)
        v = (rows == cols).to(torch.float32) - self.low.val
        keep = v != 0
        idx = torch.stack([rows[keep] - self.row0, cols[keep]])
        return torch.sparse_coo_tensor(idx, v[keep], (self.low.n_rows, self.low.n_cols)).coalesce()


def _validate_high(row, col, val, adj_high, n):
    """``adj_high`` must be ``I - adj_low`` (ACM-Pytorch/utils.py:627; Geometric train.py:78):
    the kernels derive the high-pass channel from A_low and ignore this tensor numerically."""
    hr, hc, hv = _coo_parts(adj_high)
    exp = (row == col).to(torch.float32) - val
    keep = exp != 0
    # the fp64-then-cast recipe (Geometric) may differ from fp32 I - w in the last ulp and
    # in which exact zeros were dropped -> compare densely scattered values with a tolerance
    key_e = row[keep] * n + col[keep]
    key_h = hr * n + hc
    ok = False
    if key_e.shape == key_h.shape and bool((key_e == key_h).all()):
        ok = bool(torch.allclose(exp[keep], hv, rtol=1e-5, atol=1e-6))
    else:
        allk = torch.unique(torch.cat([key_e, key_h]))
        a = torch.zeros(allk.shape[0], device=row.device)
        b = torch.zeros(allk.shape[0], device=row.device)
        a[torch.searchsorted(allk, key_e)] = exp[keep]
        b[torch.searchsorted(allk, key_h)] = hv
        ok = bool(torch.allclose(a, b, rtol=1e-5, atol=1e-6))
    if not ok:
        raise ValueError("adj_high is not I - adj_low: the ACM B200 layer derives the high-pass channel "
                         "from adj_low (A_high Z = Z - A_low Z) and cannot honour an unrelated adj_high")


# -- cache: the driver passes the same tensor objects every epoch ---------------------------
# LRU of converted operators.  An eviction only drops the cache's own reference: whoever still holds
# the AcmOperator (a layer call in flight, a GraphedTrainStep whose captured graph has the CSR
# pointers baked in) keeps its device arrays alive.
_CACHE = collections.OrderedDict()
_CACHE_MAX = 16


def _key(t):
    if t is None:
        return None
    if t.layout == torch.strided:
        return ("d", t.data_ptr(), t._version, tuple(t.shape))
    return ("s", id(t), t._version, tuple(t.shape), t._nnz())


def cached_operator(adj_low, adj_high, adj_low_unnormalized) -> AcmOperator:
    k = (_key(adj_low), _key(adj_high), _key(adj_low_unnormalized))
    hit = _CACHE.get(k)
    if hit is not None:
        op, refs = hit
        if all(r() is t for r, t in zip(refs, (adj_low, adj_high, adj_low_unnormalized)) if t is not None):
            _CACHE.move_to_end(k)
            return op
    op = AcmOperator.from_adjacency(adj_low, adj_high, adj_low_unnormalized)
    refs = tuple(weakref.ref(t) if t is not None else (lambda: None) for t in (adj_low, adj_high, adj_low_unnormalized))
    _CACHE[k] = (op, refs)
    _CACHE.move_to_end(k)
    while len(_CACHE) > _CACHE_MAX:
        _CACHE.popitem(last=False)     # least recently used
    return op
