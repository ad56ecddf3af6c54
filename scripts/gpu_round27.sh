#!/bin/bash
# r1f, call 3: tests + headline bench after the mix_bwd ring template and the degree-sorted row order
set -u
mkdir -p gpurun_out
echo "=== pytest gpu" ; timeout 420 python -m pytest tests -q -m gpu --timeout 150 > gpurun_out/pytest_gpu_r27.log 2>&1 ; echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_gpu_r27.log; grep -E "FAILED|ERROR" gpurun_out/pytest_gpu_r27.log | head -20
echo "=== bench 10M default" ; timeout 330 python bench.py --steps 6 --no-cpu-baseline --no-e2e > gpurun_out/bench_10m_r27.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_10m_r27.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['peak_mem_gb'], d['loss'], d['clocks']); print(d['roofline']['frac']); print(d['kernel_ms_per_step']); print('north', d['north_star_order']['ms_per_step'], d['north_star_order']['roofline']['frac'])"
echo "=== bench 10M row order off" ; ACMB200_ROW_ORDER=0 timeout 330 python bench.py --steps 6 --no-cpu-baseline --no-e2e > gpurun_out/bench_10m_r27_noorder.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_10m_r27_noorder.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['clocks']); print(d['kernel_ms_per_step'])"
