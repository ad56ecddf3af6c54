"""N>1 host logic on CPU: world_size-2 gloo.  The CUDA kernels cannot run here, so the
per-rank compute is stood in by the oracle's arithmetic (tests may use the oracle); what is
under test is the row-partition bookkeeping of acm_gnn_b200/dist.py -- row bounds, padding
of the last rank, the all-gather layout (row g of the gathered table = global node g),
global column ids in the per-rank CSR slices, and the gradient all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from acm_gnn_b200.dist import RowPartition
    part = RowPartition(n)
    torch.manual_seed(0)  # replicated inputs/parameters: every rank seeds identically
    fin, f = 6, 5
    row, col = O.synthetic_edges(n, 8 * n, seed=1)
    op = O.build_operator(row, col, n)
    x = torch.rand(n, fin)
    w = torch.rand(fin, 2 * f, requires_grad=True)
    # ---- full computation (reference for the partitioned one) ----
    low, _ = O.operator_to_torch(op)
    h = x @ w
    full = torch.sparse.mm(low, h)
    full.square().sum().backward()
    gw_full = w.grad.clone()
    # ---- row-partitioned: local rows of H, all-gather, local rows of the CSR ----
    r0, r1 = part.r0, part.r1
    assert (r0, r1) == part.bounds(rank) and part.n_local == r1 - r0
    w2 = w.detach().clone().requires_grad_(True)
    h_loc = x[r0:r1] @ w2
    table = part.all_gather_rows(h_loc.detach())
    assert table.shape[0] == world * part.rows_per_rank
    assert torch.allclose(table[:n], h.detach())          # row g == global node g
    e0, e1 = op.rowptr[r0], op.rowptr[r1]
    rows_loc = torch.from_numpy(op.rows()[e0:e1] - r0)
    cols = torch.from_numpy(op.col[e0:e1])                  # GLOBAL column ids
    a_loc = torch.sparse_coo_tensor(torch.stack([rows_loc, cols]), torch.from_numpy(op.w_low[e0:e1]), (r1 - r0, table.shape[0]))
    out_loc = torch.sparse.mm(a_loc, table)
    assert torch.allclose(out_loc, full[r0:r1].detach(), atol=1e-6)
    # backward: dS rows -> all-gather -> transposed aggregation of the local rows -> dW all-reduce
    ds_loc = 2 * out_loc
    ds_all = part.all_gather_rows(ds_loc)[:n]
    dh_loc = torch.sparse.mm(low.t(), ds_all)[r0:r1]
    gw = x[r0:r1].t() @ dh_loc
    part.all_reduce_(gw)
    assert torch.allclose(gw, gw_full, rtol=1e-4, atol=1e-6)
    # ---- protocol decisions must be GLOBAL: only rank 0's slice has a long row -> every rank must refuse the
    # rank-structured backward table (functional.use_rank1_table); with no long row anywhere every rank accepts it
    from acm_gnn_b200.functional import LayerConfig, use_rank1_table

    class _Csr:
        def __init__(self, lr):
            self.col, self._lr = torch.zeros(1, dtype=torch.int32), lr

        def long_rows(self, transposed=False):
            return self._lr

    class _Op:
        def __init__(self, lr):
            self.low = _Csr(lr)

    cfg = LayerConfig(variant=True, dist=part)
    os.environ.pop("ACMB200_BWD_RANK1", None)
    assert use_rank1_table(cfg, _Op(("rows",) if rank == 0 else None), 256) is False
    assert use_rank1_table(cfg, _Op(None), 256) is True
    # ---- replicated parameters outside the ACM layers (the mlpX branch of acmgcn++) see only the local rows:
    # dist.attach hooks their gradients into the all-reduce; the result equals the full-batch gradient
    from acm_gnn_b200.dist import attach
    torch.manual_seed(3)
    full_m, part_m = torch.nn.Linear(fin, 4), torch.nn.Linear(fin, 4)
    part_m.load_state_dict(full_m.state_dict())
    full_m(x).relu().sum().backward()
    attach(part_m, part)
    attach(part_m, part)                                   # idempotent: re-attaching must not reduce twice
    part_m(x[r0:r1]).relu().sum().backward()
    assert torch.allclose(part_m.weight.grad, full_m.weight.grad, rtol=1e-5, atol=1e-6)
    assert torch.allclose(part_m.bias.grad, full_m.bias.grad, rtol=1e-5, atol=1e-6)
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([1]))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [64, 77])  # divisible and ragged (last rank padded)
def test_row_partition_world2_gloo(tmp_path, n):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0.npy") and os.path.exists(tmp_path / "ok1.npy")


def test_partition_bounds_cover_all_rows():
    from acm_gnn_b200.dist import RowPartition

    class Fake(RowPartition):
        def __init__(self, n, world, rank):
            self.world, self.rank, self.n_global = world, rank, n
            self.rows_per_rank = (n + world - 1) // world
            self.r0 = min(n, rank * self.rows_per_rank)
            self.r1 = min(n, self.r0 + self.rows_per_rank)
            self.n_local = self.r1 - self.r0

    for n, world in ((10, 8), (10_000_000, 8), (7, 2), (5, 8)):
        covered = []
        for r in range(world):
            p = Fake(n, world, r)
            covered += list(range(p.r0, p.r1)) if n < 100 else [p.r1 - p.r0]
        if n < 100:
            assert covered == list(range(n))
        else:
            assert sum(covered) == n
