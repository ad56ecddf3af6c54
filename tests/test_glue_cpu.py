"""CPU side of the inter-layer glue (acm_glue_fwd / acm_glue_bwd): the numpy restatement of its Philox4x32-10
mask is pinned by the generator's published known-answer vectors, and the host mirror keeps the reference's own
torch ops wherever the fused launch would not be bit-identical to them."""
import numpy as np
import torch

from helpers import ROOT  # noqa: F401  (puts the repo root on sys.path)
from oracle import philox_oracle as P


def test_philox_known_answers():
    for ctr, key, out in P.KAT:
        got = P.philox4x32_10(*ctr, *key)
        assert tuple(int(v) for v in got) == out, (ctr, key)


def test_keep_bits_rate_layout_and_offsets():
    total = 1_000_003
    a = P.keep_bits(1234, 0, total, 0.3)
    b = P.keep_bits(1234, 1, total, 0.3)
    c = P.keep_bits(1235, 0, total, 0.3)
    assert abs(a.mean() - 0.7) < 3e-3 and abs(b.mean() - 0.7) < 3e-3
    assert 0.4 < (a == b).mean() < 0.7 and 0.4 < (a == c).mean() < 0.7       # independent masks: agreement 0.58
    assert P.keep_bits(1234, 0, 1000, 0.3).tolist() == a[:1000].tolist()      # a prefix does not depend on the length
    assert P.keep_bits(7, 0, 100, 0.0).all()
    packed = P.pack_mask(a)
    assert packed.shape[0] == (total + 7) // 8
    assert bool(packed[0] >> 3 & 1) == bool(a[3]) and bool(packed[5] >> 7 & 1) == bool(a[47])
    assert P.dropout_threshold(0.5) == 1 << 31 and P.dropout_threshold(0.0) == 0


def test_oracle_glue_equals_reference_ops_without_dropout():
    g = torch.Generator().manual_seed(0)
    x, add = torch.randn(1001, 37, generator=g), torch.randn(1001, 37, generator=g)
    for dt in (torch.float32, torch.bfloat16):
        for relu in (False, True):
            y, on = P.glue_forward(x.to(dt), add.to(dt), relu, 0.0, 1, 0)
            ref = (torch.relu(x.to(dt)) if relu else x.to(dt)) + add.to(dt)
            assert torch.equal(y, ref)
            assert bool(on.all()) == (not relu)
    # bf16 activations + fp32 xX: torch promotes the sum to fp32
    y, _ = P.glue_forward(x.bfloat16(), add, True, 0.0, 1, 0)
    assert y.dtype == torch.float32 and torch.equal(y, torch.relu(x.bfloat16()) + add)


def test_host_mirror_keeps_reference_ops_when_not_bit_identical(monkeypatch):
    """p > 0 in training: the reference's own F.relu / F.dropout / + run unless ACMB200_FUSED_DROPOUT=1 (checked
    on CPU tensors, where the fused launch is never eligible: same generator -> same values as the reference line)."""
    from acm_gnn_b200.functional import inter_layer_glue
    x, add = torch.randn(64, 16), torch.randn(64, 16)
    torch.manual_seed(5)
    ref = torch.nn.functional.dropout(torch.relu(x), 0.4, training=True) + add
    torch.manual_seed(5)
    assert torch.equal(inter_layer_glue(x, add, True, 0.4, True), ref)
    assert torch.equal(inter_layer_glue(x, None, True, 0.4, False), torch.relu(x))
    assert inter_layer_glue(x, None, False, 0.4, False) is x                  # identity: no launch at all
