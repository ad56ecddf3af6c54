"""CPU restatement of the inter-layer glue kernel (acm_glue_fwd / acm_glue_bwd)  --  TEST INFRASTRUCTURE ONLY.

Only ``tests/`` may import this module; the product path never does.

What it restates.  The reference's glue between its two layers is
``fea1 = F.dropout(F.relu(fea1), p, training) [+ xX]`` (ACM-Pytorch/models/models.py:160-164,
ACM-Geometric/models.py:70-74).  Its relu / add arithmetic is deterministic and is checked against torch
directly.  The dropout mask is random: the reference takes it from torch's generator, whose stream
depends on the launch geometry of torch's own kernel and cannot be (and is not claimed to be) reproduced.
The product kernel instead draws bits from Philox4x32-10 with a documented counter layout
(include/acm_b200.h: element e is kept iff philox(seed, [e // 4, offset])[e % 4] >= floor(p * 2^32));
this file restates exactly that in numpy so the GPU mask can be compared BIT FOR BIT.

Parity status: the generator is pinned by the published known-answer vectors of Philox4x32-10
(Random123 ``kat_vectors``: Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3",
SC'11), see ``KAT`` below and tests/test_glue_cpu.py; the dropout SEMANTICS (keep with probability 1 - p,
scale kept values by 1 / (1 - p), ``torch.nn.functional.dropout``) are checked statistically and through
the identity ``y == F.relu(x) * mask / (1 - p) + add``.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)

# (counter, key) -> output, from Random123's kat_vectors for philox4x32 with 10 rounds
KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10: uint32 arrays (or scalars) in, four uint32 arrays out."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32).copy() for c in (c0, c1, c2, c3))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK32).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0, k1 = np.uint32(k0 + W0), np.uint32(k1 + W1)
    return c0, c1, c2, c3


def dropout_threshold(p: float) -> int:
    """floor(p * 2^32) clamped to 2^32 - 1, with p taken as the fp32 value the C ABI receives."""
    t = float(np.float32(p)) * 4294967296.0
    return 0xFFFFFFFF if t >= 4294967295.0 else int(t)


def keep_bits(seed: int, offset: int, total: int, p: float) -> np.ndarray:
    """bool[total]: element e survives the dropout."""
    thr = dropout_threshold(p)
    if thr == 0:
        return np.ones(total, dtype=bool)
    g = np.arange((total + 3) // 4, dtype=np.uint64)
    seed, offset = int(seed) & (2 ** 64 - 1), int(offset) & (2 ** 64 - 1)
    r = philox4x32_10((g & MASK32).astype(np.uint32), (g >> np.uint64(32)).astype(np.uint32),
                      np.full(g.shape, offset & 0xFFFFFFFF, dtype=np.uint32), np.full(g.shape, offset >> 32, dtype=np.uint32),
                      seed & 0xFFFFFFFF, seed >> 32)
    return (np.stack(r, axis=1).reshape(-1)[:total] >= np.uint32(thr))


def pack_mask(on: np.ndarray) -> np.ndarray:
    """bit j of byte t = element 8 t + j (the layout acm_glue_fwd writes)."""
    return np.packbits(on.astype(np.uint8), bitorder="little")


def glue_forward(x, add, relu: bool, p: float, seed: int, offset: int):
    """(y, pass-mask) of the glue on torch CPU tensors, in the op order of the reference (relu, dropout
    scaling rounded to the storage dtype, then the add)."""
    import torch
    keep = torch.from_numpy(keep_bits(seed, offset, x.numel(), p)).view(x.shape)
    on = keep & (x > 0) if relu else keep
    scale = float(np.float32(1.0) / (np.float32(1.0) - np.float32(p))) if dropout_threshold(p) else 1.0
    y = torch.where(on, (x.float() * scale).to(x.dtype), torch.zeros((), dtype=x.dtype))
    if add is not None:
        y = y + add
    return y, on
