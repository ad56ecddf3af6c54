#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== tc probe (persistent tn kernel)" ; timeout 300 python scripts/tc_probe.py > gpurun_out/tc_probe.log 2>&1 ; echo "tc_probe rc=$?" ; tail -9 gpurun_out/tc_probe.log
echo "=== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -5 gpurun_out/pytest_gpu.log
echo "=== bench 10M" ; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_10m_r6.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_10m_r6.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e'], d['kernel_ms_per_step'])"
