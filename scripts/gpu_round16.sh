#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest gpu all" ; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -6 gpurun_out/pytest_gpu.log
echo "=== smoke" ; timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -3 gpurun_out/smoke.log
for G in 1 0; do
echo "=== bench 10M gather=$G" ; ACMB200_GATHER=$G timeout 900 python bench.py --steps 6 --no-cpu-baseline --no-e2e > gpurun_out/bench_10m_gather$G.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_10m_gather$G.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); n=d['north_star_order']; print(d['ms_per_step'], d['value'], d['roofline']['frac'], 'north', n['ms_per_step'], n['roofline']['avg_launch_ms'], n['roofline']['frac'], d['kernel_ms_per_step'])"
done
echo "=== bench 10M fp32-boundary activations" ; ACMB200_BF16_ACT=0 timeout 900 python bench.py --steps 6 --no-cpu-baseline --no-e2e > gpurun_out/bench_10m_act0.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_10m_act0.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['kernel_ms_per_step'])"
