"""bench.py contract on CPU: the reference arm (--impl reference) runs without a GPU, prints ONE
JSON line with the agreed keys, and non-zero ranks of a torchrun launch exit without work."""
import json
import os
import subprocess
import sys

from helpers import ROOT


def _run(extra_env=None, args=()):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", **(extra_env or {}))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--nodes", "20000", "--edges", "400000", "--cpu-nodes", "2000",
                        "--fin", "32", "--hidden", "32", "--nclass", "4", *args],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_reference_arm_prints_one_json_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "edges/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    staged = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "ACM-Pytorch", "models", "models.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port")
    assert d["cpu_baseline"]["cores"] == os.cpu_count()
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["metric"] == "acm_gcn_train_step_edges_per_sec"


def test_reference_arm_nonzero_rank_is_silent():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, ("--gpus", "2")) == []


def test_reference_arm_falls_back_to_the_oracle_port():
    """Without the staged reference tree (forced here) the CPU arm times oracle/acm_oracle.py."""
    d = json.loads(_run({"ACMB200_BENCH_PORT": "1"})[0])
    assert d["cpu_baseline"]["kind"] == "port" and d["value"] > 0
