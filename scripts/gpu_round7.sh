#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest gpu (aggregate-first)" ; timeout 900 python -m pytest tests -m gpu -q -k "aggregate_first" > gpurun_out/pytest_gpu7.log 2>&1 ; echo "pytest rc=$?" ; tail -15 gpurun_out/pytest_gpu7.log
echo "=== pytest gpu all" ; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -4 gpurun_out/pytest_gpu.log
for R in auto off; do
echo "=== bench 10M reorder=$R" ; timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --reorder $R > gpurun_out/bench_10m_reorder_$R.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_10m_reorder_$R.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline'], d['e2e']['value'], d['peak_mem_gb'], d['loss'], d['kernel_ms_per_step'])"
done
