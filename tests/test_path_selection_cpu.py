"""Host logic that decides WHICH kernels a layer call runs (acm_gnn_b200/functional.py predicates) and the bench presets:
pure Python, no GPU.  The kernels behind every path are parity-tested on the GPU (tests/test_gpu_parity.py); these
tests pin the selection rules and the environment knobs documented in DESIGN.md section 3."""
import json
import os
import subprocess
import sys

import pytest

from helpers import ROOT


class _Csr:
    def __init__(self, long_rows=None):
        import torch
        self._lr = long_rows
        self.col = torch.zeros(1, dtype=torch.int32)       # device of the flag the ranks all-reduce

    def long_rows(self, transposed=False):
        return self._lr


class _Op:
    def __init__(self, long_rows=None):
        self.low = _Csr(long_rows)


@pytest.fixture(autouse=True)
def _clean_env(monkeypatch):
    for k in ("ACMB200_REORDER", "ACMB200_BWD_INPUT", "ACMB200_LOCAL_TABLE", "ACMB200_BWD_RANK1", "ACMB200_FUSED_FWD", "ACMB200_DTYPE"):
        monkeypatch.delenv(k, raising=False)


def test_default_dtype_is_the_reference_precision(monkeypatch):
    from acm_gnn_b200.functional import default_dtype
    assert default_dtype() == "fp32"                 # the drop-in computes like the reference unless the user opts in
    monkeypatch.setenv("ACMB200_DTYPE", "bf16")
    assert default_dtype() == "bf16"
    monkeypatch.setenv("ACMB200_DTYPE", "fp8")
    with pytest.raises(ValueError):
        default_dtype()


def test_aggregate_first_and_input_backward_rules(monkeypatch):
    from acm_gnn_b200.functional import LayerConfig, use_aggregate_first, use_input_backward
    cfg = LayerConfig(variant=False, dtype="bf16")
    assert use_aggregate_first(cfg, 256, 256, x_needs_grad=False)
    assert not use_aggregate_first(cfg, 256, 256, x_needs_grad=True)        # dX needs the transposed aggregation
    assert not use_aggregate_first(cfg, 1433, 64, x_needs_grad=False)       # Cora: input row wider than the table row
    assert not use_aggregate_first(LayerConfig(variant=True), 64, 64, False)  # relu before the aggregation
    monkeypatch.setenv("ACMB200_REORDER", "off")
    assert not use_aggregate_first(cfg, 256, 256, False)
    # transform-first order: the backward still aggregates the INPUT when the layer input carries no gradient
    assert use_input_backward(cfg, 256, 256, x_needs_grad=False)
    assert not use_input_backward(cfg, 256, 256, x_needs_grad=True)
    assert not use_input_backward(LayerConfig(variant=True), 256, 256, False)
    assert not use_input_backward(cfg, 2089, 64, False)                     # Squirrel layer 0
    monkeypatch.setenv("ACMB200_BWD_INPUT", "off")
    assert not use_input_backward(cfg, 256, 256, False)
    monkeypatch.setenv("ACMB200_BWD_INPUT", "maybe")
    with pytest.raises(ValueError):
        use_input_backward(cfg, 256, 256, False)


def test_partition_exchanges_the_narrower_operand(monkeypatch):
    from acm_gnn_b200.functional import LayerConfig, use_local_table, use_rank1_table
    class _Part:                      # stands in for dist.RowPartition: the other ranks report no long rows either
        def all_reduce_(self, t):
            return t
    part = _Part()
    cfg = LayerConfig(dist=part, dtype="bf16")
    assert use_local_table(cfg, ldx=256, fp=256)          # X (256 wide) instead of [HL|HH] (512 wide)
    assert not use_local_table(cfg, ldx=256, fp=16)       # layer 1: the 32-wide table is the narrower one
    assert not use_local_table(LayerConfig(dist=None), 256, 256)
    monkeypatch.setenv("ACMB200_LOCAL_TABLE", "off")
    assert not use_local_table(cfg, 256, 256)
    # rank-structured backward table: variant 1 without LayerNorm, wide rows, no long rows; default only under a partition
    v1 = LayerConfig(variant=True, dist=part)
    assert use_rank1_table(v1, _Op(), 256)
    assert not use_rank1_table(LayerConfig(variant=True, dist=None), _Op(), 256)
    assert not use_rank1_table(LayerConfig(variant=False, dist=part), _Op(), 256)
    assert not use_rank1_table(LayerConfig(variant=True, ln_live=True, dist=part), _Op(), 256)
    assert not use_rank1_table(v1, _Op(), 16)
    assert not use_rank1_table(v1, _Op(long_rows=("rows",)), 256)
    monkeypatch.setenv("ACMB200_BWD_RANK1", "on")
    assert use_rank1_table(LayerConfig(variant=True, dist=None), _Op(), 256)
    monkeypatch.setenv("ACMB200_BWD_RANK1", "off")
    assert not use_rank1_table(v1, _Op(), 256)


def test_fused_forward_rule(monkeypatch):
    from acm_gnn_b200 import _lib
    from acm_gnn_b200.functional import LayerConfig, use_fused_forward
    cfg = LayerConfig(dtype="bf16", out_dtype="bf16")
    tc, simt = _lib.GEMM_TCGEN05, _lib.GEMM_SIMT
    assert use_fused_forward(cfg, tc, 256, 256, 3, 256)
    assert not use_fused_forward(cfg, simt, 256, 256, 3, 256)
    assert not use_fused_forward(cfg, tc, 64, 64, 3, 256)                     # only the 256-wide layer
    assert not use_fused_forward(cfg, tc, 256, 256, 4, 256)                   # structure channel
    assert not use_fused_forward(LayerConfig(dtype="bf16", ln_live=True), tc, 256, 256, 3, 256)
    assert not use_fused_forward(cfg, tc, 256, 200, 3, 256)                   # bf16 Y rows must be 32-byte multiples
    assert use_fused_forward(LayerConfig(dtype="bf16", out_dtype="fp32"), tc, 256, 200, 3, 256)
    monkeypatch.setenv("ACMB200_FUSED_FWD", "off")
    assert not use_fused_forward(cfg, tc, 256, 256, 3, 256)


def test_out_features_limit_fails_at_construction():
    import acm_gnn_b200.layers as L
    with pytest.raises(NotImplementedError):
        L.GraphConvolution(16, 300, 10, "acmgcn")


def test_bench_presets_and_shared_config():
    """--config cfg3 / cfg4 fill the BASELINE shapes; both arms print the SAME config dict (incl. what the CPU arm samples)."""
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    old = sys.argv
    try:
        sys.argv = ["bench.py", "--config", "cfg3"]
        a = bench.parse_args()
        assert (a.nodes, a.edges, a.fin, a.hidden, a.nclass) == (168_114, 13_595_114, 7, 256, 2)
        assert (a.model_type, a.variant, a.flavour, a.features) == ("acmgcnp", 1, "geometric", "normal")
        sys.argv = ["bench.py", "--config", "cfg4", "--hidden", "128"]
        b = bench.parse_args()
        assert b.model_type == "acmgcnpp" and b.hidden == 128 and b.fin == 128        # explicit flags override the preset
        sys.argv = ["bench.py"]
        c = bench.parse_args()
        assert (c.nodes, c.edges, c.model_type, c.variant) == (10_000_000, 200_000_000, "acmgcn", 0)
        cfg = bench.workload_config(c, 1)
        assert "N=100000" in cfg["reference_arm_sample"] and cfg["preset"] == "cfg5" and "cfg 5" in cfg["workload"]
        assert json.dumps(cfg) == json.dumps(bench.workload_config(c, 1))
    finally:
        sys.argv = old


def test_reference_arm_runs_the_geometric_preset_on_cpu():
    """cfg3 through the reference's own ACM-Geometric/models.py (staged) or the oracle port: one JSON line, config says what ran."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "cfg3", "--steps", "1",
                        "--warmup", "0", "--cpu-nodes", "1500"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert d["impl"] == "reference" and d["value"] > 0 and d["sample_nodes"] == 1500
    assert "N=1500" in d["config"]["reference_arm_sample"] and "geometric" in d["config"]["workload"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port")
