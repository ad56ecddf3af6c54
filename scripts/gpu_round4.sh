#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest gpu (new tests)" ; timeout 900 python -m pytest tests -m gpu -q -k "nll or train_py" > gpurun_out/pytest_gpu4.log 2>&1 ; echo "pytest rc=$?" ; tail -15 gpurun_out/pytest_gpu4.log
echo "=== tc probe" ; timeout 300 python scripts/tc_probe.py > gpurun_out/tc_probe.log 2>&1 ; echo "tc_probe rc=$?" ; tail -2 gpurun_out/tc_probe.log
for L2 in 0 32 64 128; do
echo "=== bench 10M fused loss l2=$L2" ; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --l2-fetch $L2 > gpurun_out/bench_10m_l2_$L2.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_10m_l2_$L2.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['kernel_ms_per_step'])"
done
