#!/bin/bash
# r1f, call 6: tests (+ bench contract tests), agg-first ring vs LDG A/B
set -u
mkdir -p gpurun_out
echo "=== pytest gpu" ; timeout 500 python -m pytest tests -q -m gpu --timeout 200 > gpurun_out/pytest_gpu_r29.log 2>&1 ; echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_gpu_r29.log; grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu_r29.log | head -20
SUM='
import json,sys
d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["clocks"]["sm_mhz"], "roof", d["roofline"]["frac"], d["roofline"]["avg_launch_ms"]); print(d["kernel_ms_per_step"])'
echo "=== bench 10M (agg-first ring)" ; timeout 330 python bench.py --steps 8 --no-cpu-baseline --no-e2e > gpurun_out/bench_10m_r29_ring.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_10m_r29_ring.log | python -c "$SUM"
echo "=== bench 10M (ACMB200_GATHER=0: LDG)" ; ACMB200_GATHER=0 timeout 330 python bench.py --steps 8 --no-cpu-baseline --no-e2e > gpurun_out/bench_10m_r29_ldg.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_10m_r29_ldg.log | python -c "$SUM"
