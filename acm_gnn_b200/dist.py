"""1-D row (node) partition of the ACM layer over the GPUs of one box.

The reference is single-device.  Rows of ``A_low`` (destination nodes) are independent;
each output row needs the table rows of its neighbours, so the one data-path exchange per
aggregation is an all-gather of the operand table -- ``[HL|HH]`` forward, ``[dS_L|dS_H]``
backward -- over NVLink (NCCL), plus an all-reduce of the (replicated) parameter
gradients.  For uniform random graphs the halo is ~all nodes (each of 8 ranks references
92 % of them at mean degree 20), so a full all-gather is the right collective.

One process per GPU (torchrun); rank r owns rows [r*rows_per_rank, min(N,(r+1)*rows_per_rank)).
Works with the gloo backend on CPU tensors too, which is how the host logic is tested.
"""
from __future__ import annotations

import ctypes
import os

import torch
import torch.distributed as dist


class RowPartition:
    def __init__(self, n_global: int, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n_global = int(n_global)
        self.rows_per_rank = (self.n_global + self.world - 1) // self.world
        self.r0 = min(self.n_global, self.rank * self.rows_per_rank)
        self.r1 = min(self.n_global, self.r0 + self.rows_per_rank)
        self.n_local = self.r1 - self.r0
        self.bytes_gathered = 0
        self._symm = {}
        self._push_ok = None
        self.multicast = False   # set by symm_table: exchange pushes go through NVSwitch multicast

    def bounds(self, rank):
        r0 = min(self.n_global, rank * self.rows_per_rank)
        return r0, min(self.n_global, r0 + self.rows_per_rank)

    def all_gather_rows(self, local: torch.Tensor) -> torch.Tensor:
        """[n_local, W] per rank -> [world*rows_per_rank, W] on every rank; row g of the result
        is global node g (ranks are padded to rows_per_rank rows; pad rows are never indexed
        because column ids are < n_global)."""
        w = local.shape[1]
        if local.shape[0] != self.n_local:
            raise ValueError(f"expected {self.n_local} local rows, got {local.shape[0]}")
        if self.n_local != self.rows_per_rank:
            pad = torch.zeros(self.rows_per_rank, w, dtype=local.dtype, device=local.device)
            pad[: self.n_local] = local
            local = pad
        out = torch.empty(self.world * self.rows_per_rank, w, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=self.group)
        self.bytes_gathered += out.numel() * out.element_size()
        return out

    # -- fused exchange: tables in symmetric (peer-mapped) memory ---------------------------------
    def push_enabled(self) -> bool:
        """Kernels store operand-table rows straight into every rank's table over NVLink peer
        mappings (csrc PeerTables) instead of a separate NCCL all-gather.  Needs CUDA symmetric
        memory across the ranks of one box (<= 8); ACMB200_PUSH=0 forces the NCCL path."""
        if self._push_ok is None:
            ok = os.environ.get("ACMB200_PUSH", "1") != "0" and self.world <= 8 and torch.cuda.is_available()
            if ok:
                try:  # probe once: a tiny symmetric allocation + rendezvous (collective over the group)
                    import torch.distributed._symmetric_memory as symm
                    torch.cuda.synchronize()            # rendezvous with drained devices and aligned hosts (see symm_table)
                    dist.barrier(group=self.group)
                    t = symm.empty((1024,), dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
                    hdl = symm.rendezvous(t, self.group if self.group is not None else dist.group.WORLD)
                    ok = len(hdl.buffer_ptrs) == self.world
                    hdl.barrier(channel=0)
                    self._probe = (t, hdl)
                except Exception as e:  # NCCL all-gather remains the exchange
                    import warnings
                    warnings.warn(f"acm_gnn_b200: symmetric memory unavailable ({e!r}); using NCCL all-gather")
                    ok = False
            self._push_ok = ok
        return self._push_ok

    def symm_table(self, key, width, dtype, device):
        """Persistent [world*rows_per_rank, width] table in symmetric memory, TWO per (layer,
        direction), used alternately.  Returns (tensor, handle, ctypes array of the ``world``
        peer base pointers, multicast address or 0).  The multicast (NVLS) address is non-zero when
        the driver mapped the allocation into an NVSwitch multicast object: the kernels then issue
        one ``multimem.st`` per 16 bytes instead of ``world`` peer stores.  Opt-in with
        ACMB200_MULTICAST=1: measured on 4 B200s it is slightly SLOWER than the unicast stores (an
        all-gather is bound by every rank's NVLink ingress, and the multicast loop-back of the own
        rows adds 1/world to it), so unicast stays the default.

        Alternating buffers removes the "everybody is done reading" barrier before a push: a
        buffer is rewritten two uses later, and by then every rank has passed the post-push
        barrier of the use in between, which it can only reach after finishing its reads of this
        buffer (stream order).  Consequence: at most TWO forwards of the same layer may be
        outstanding before their backwards run (variant 1 keeps a view of the forward table for
        its relu mask); deeper gradient accumulation needs ACMB200_PUSH=0."""
        par = self._symm.get(("parity", key), 0)
        self._symm[("parity", key)] = par ^ 1
        k = (key, par, width, dtype)
        hit = self._symm.get(k)
        if hit is not None:
            return hit
        import torch.distributed._symmetric_memory as symm
        # First use of this table (warm-up steps only): allocate + rendezvous with IDLE devices and aligned hosts.
        # The rendezvous maps peer memory and talks to the store while holding the GIL; doing that while this GPU
        # still runs a kernel that waits for a peer (exchange barrier, NCCL) whose host is itself blocked in a later
        # rendezvous can deadlock (seen once in ~15 two-GPU runs: rank 0 stuck before its all-reduce, rank 1 in the
        # next rendezvous).  Draining the device and meeting at a barrier first removes every such interleaving.
        torch.cuda.synchronize(device)
        dist.barrier(group=self.group)
        t = symm.empty((self.world * self.rows_per_rank, width), dtype=dtype, device=device)
        hdl = symm.rendezvous(t, self.group if self.group is not None else dist.group.WORLD)
        torch.cuda.synchronize(device)
        ptrs = (ctypes.c_void_p * self.world)(*[int(p) for p in hdl.buffer_ptrs])
        mc = 0
        if os.environ.get("ACMB200_MULTICAST", "0") == "1":
            mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
            if mc:  # same offset inside the allocation as the tensor has from its unicast base
                mc += t.data_ptr() - int(hdl.buffer_ptrs[dist.get_rank(self.group)])
        self.multicast = bool(mc)
        self._symm[k] = (t, hdl, ptrs, mc)
        return self._symm[k]

    def all_reduce_(self, t: torch.Tensor) -> torch.Tensor:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t


def attach(model, part: RowPartition):
    """Mark every ACM layer of ``model`` as row-partitioned (parameters stay replicated: all
    ranks seed identically, SURVEY.md 8e "Process model")."""
    from .layers import GraphConvolution
    own = set()
    for m in model.modules():
        if isinstance(m, GraphConvolution):
            m.acm_dist = part
            own.update(id(p) for p in m.parameters(recurse=False))
    # Replicated parameters OUTSIDE the ACM layers (the mlpX branch of acmgcn++) see only this rank's rows: their
    # gradients are partial sums like the layers' own (which AcmLayerFunction all-reduces itself) -> all-reduce them as
    # autograd produces them.  (BatchNorm statistics of a deeper mlpX stay per rank.)
    for h in getattr(model, "_acm_dist_hooks", []):
        h.remove()
    model._acm_dist_hooks = [p.register_hook(lambda g, part=part: part.all_reduce_(g.contiguous().clone()))
                             for p in model.parameters() if id(p) not in own and p.requires_grad]
    return model
