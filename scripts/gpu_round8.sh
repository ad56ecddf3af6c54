#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== tc probe (smem-staged epilogue)" ; timeout 300 python scripts/tc_probe.py > gpurun_out/tc_probe.log 2>&1 ; echo "tc_probe rc=$?" ; tail -3 gpurun_out/tc_probe.log
echo "=== pytest gpu all" ; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -4 gpurun_out/pytest_gpu.log
echo "=== bench 10M default" ; timeout 900 python bench.py > gpurun_out/bench_10m_r8.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_10m_r8.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['north_star_order'], d['kernel_ms_per_step'])"
