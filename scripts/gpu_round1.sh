#!/bin/bash
# First GPU pass: smoke, parity tests, small + full bench.  Logs under gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== smoke" ; timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -5 gpurun_out/smoke.log
echo "=== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -25 gpurun_out/pytest_gpu.log
echo "=== bench 1M" ; timeout 900 python bench.py --nodes 1000000 --edges 20000000 --steps 5 --warmup 3 > gpurun_out/bench_1m.log 2>&1 ; echo "bench1m rc=$?" ; tail -2 gpurun_out/bench_1m.log
