"""Multi-GPU parity check (run under torchrun, one rank per GPU): the row-partitioned model
(NCCL all-gather of the operand tables, all-reduce of parameter gradients) must reproduce the
single-GPU model on the same graph, inputs and parameters.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dist_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    import acm_gnn_b200 as A
    from acm_gnn_b200 import layers as L
    from acm_gnn_b200.dist import RowPartition, attach
    from acm_gnn_b200.functional import nll_log_softmax
    L.device = dev
    ok_all = True
    # hidden 64: narrow-row (smem-staged) pushes; hidden 256 + variant 1: wide-row pushes in the GEMM
    # epilogue and in mix_bwd.  Layer 0 of every case has Fin = 48 <= 2*hidden, so by default its
    # [HL|HH] table is built locally from the all-gathered input (functional.use_local_table) and a
    # variant-0 layer 0 runs its backward through the input aggregation (use_input_backward);
    # the cases with LOCAL_TABLE / BWD_INPUT = off keep the fused table push of both directions covered.
    OLD = {"ACMB200_LOCAL_TABLE": "off", "ACMB200_BWD_INPUT": "off"}
    NS = {"ACMB200_REORDER": "off"}                      # north-star order: transform-first in layer 0 too
    cases = [("fp32", False, 5003, False, 0, 64, {}), ("bf16", False, 5003, False, 0, 64, {}),
             ("fp32", True, 4096, False, 0, 64, {}), ("fp32", False, 5003, True, 0, 64, {}),
             ("bf16", False, 4100, True, 0, 64, {}), ("bf16", True, 4100, False, 0, 64, {}),
             ("fp32", False, 3001, False, 1, 64, {}), ("bf16", True, 3001, False, 1, 64, {}),
             ("bf16", True, 4100, False, 0, 256, {}), ("fp32", True, 4100, False, 0, 256, {}),
             ("bf16", False, 4100, True, 0, 256, {}),
             ("bf16", False, 4100, True, 0, 256, NS), ("fp32", False, 5003, False, 0, 64, NS),
             ("bf16", True, 4100, True, 0, 256, NS), ("bf16", False, 4100, False, 0, 256, NS),
             ("bf16", True, 4100, False, 0, 256, OLD), ("fp32", False, 5003, False, 0, 64, dict(OLD, **NS)),
             ("bf16", False, 4100, True, 0, 256, dict(OLD, **NS)), ("bf16", True, 3001, False, 1, 64, OLD),
             # rank-structured backward table (variant 1: G row + 4 scalars exchanged instead of [dO_L|dO_H]) is the
             # default under a partition; these keep the plain 2F-wide table push covered
             ("bf16", True, 4100, False, 0, 256, {"ACMB200_BWD_RANK1": "off"}), ("fp32", True, 4096, False, 0, 64, {"ACMB200_BWD_RANK1": "off"}),
             # a hub node gives ONE rank a long row (> 256 edges) of the transposed operator: the table layout is a protocol
             # between the ranks, so every rank must fall back to the plain table together (HUB marker, see below)
             ("fp32", True, 4099, False, 0, 64, {"HUB": "1"}), ("bf16", True, 4099, True, 0, 256, {"HUB": "1"})]
    if os.environ.get("ACMB200_DIST_CHECK_PP") == "1":
        # acmgcn++ (PP marker): the mlpX branch has replicated parameters OUTSIDE the ACM layers, whose gradients
        # dist.attach hooks into the all-reduce.  Opt-in: added after the last multi-GPU session of round 2, so far
        # covered by the gloo test only (tests/test_dist_cpu.py).
        cases += [("fp32", True, 4100, False, 0, 64, {"PP": "1"}), ("bf16", True, 4100, True, 0, 256, {"PP": "1"}),
                  ("bf16", False, 4100, False, 0, 256, {"PP": "1"})]
    knobs = ("ACMB200_LOCAL_TABLE", "ACMB200_BWD_INPUT", "ACMB200_REORDER", "ACMB200_BWD_RANK1")
    for mode, variant, n, staged, struct, hid, env in cases:
        for k in knobs:
            os.environ.pop(k, None)
        os.environ.update({k: v for k, v in env.items() if k.startswith("ACMB200_")})
        os.environ["ACMB200_DTYPE"] = mode
        fin, ncls = 48, 7
        g = torch.Generator(device=dev); g.manual_seed(5)
        src = torch.randint(0, n, (40000,), generator=g, device=dev)
        dst = torch.randint(0, n, (40000,), generator=g, device=dev)
        keep = src != dst
        row, col = torch.cat([src[keep], dst[keep]]), torch.cat([dst[keep], src[keep]])
        if env.get("HUB"):
            hub = torch.arange(1, 400, device=dev)
            row = torch.cat([row, torch.zeros_like(hub), hub])
            col = torch.cat([col, hub, torch.zeros_like(hub)])
        key = torch.unique(row * n + col)
        row, col = key // n, key % n
        x = torch.rand(n, fin, generator=g, device=dev)
        labels = torch.randint(0, ncls, (n,), generator=g, device=dev)
        mask = (torch.rand(n, generator=g, device=dev) < 0.6).to(torch.uint8)
        ntr = int(mask.sum().item())
        op = A.AcmOperator.from_edges(row, col, n, with_raw=bool(struct))
        mtype = "acmgcnpp" if env.get("PP") else ("acmgcnp" if struct else "acmgcn")

        def build():
            torch.manual_seed(42)
            return A.GCN(fin, hid, ncls, 2, n, 0.0, mtype, struct, variant=variant).to(dev)

        # single-GPU reference (every rank computes it redundantly)
        m1 = build()
        out1 = m1(x, op, None, None)
        nll_log_softmax(out1, labels, mask, n_train=ntr).backward()
        # row-partitioned
        part = RowPartition(n)
        m2 = attach(build(), part)
        opl = op.partition(part.r0, part.r1)
        x_loc = x[part.r0:part.r1].contiguous()
        if staged:
            x_loc = A.stage_input(x_loc, mode, part)
        out2 = m2(x_loc, opl, None, None)
        nll_log_softmax(out2, labels[part.r0:part.r1].contiguous(), mask[part.r0:part.r1].contiguous(), n_train=ntr).backward()
        torch.cuda.synchronize()
        tol = 2e-5 if mode == "fp32" else 3e-2
        e_out = float((out2 - out1[part.r0:part.r1]).abs().max() / out1.abs().max())
        worst = 0.0
        for (k, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
            if p1.grad is None or k in ("fea_param", "xX_param"):
                continue
            worst = max(worst, float((p1.grad - p2.grad).norm() / p1.grad.norm().clamp_min(1e-20)))
        ok = e_out <= tol and worst <= (2e-4 if mode == "fp32" else 0.1)
        t = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok_all = ok_all and bool(t.item())
        if rank == 0:
            print(f"dist_check world={world} mode={mode} variant={variant} staged={staged} struct={struct} hid={hid} env={env} push={part.push_enabled()} multicast={part.multicast} n={n}: out rel.err {e_out:.2e}, worst grad rel.fro {worst:.2e} -> {'OK' if t.item() else 'FAIL'}", flush=True)
    if rank == 0:
        print("DIST_CHECK", "PASS" if ok_all else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
