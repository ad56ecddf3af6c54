#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== tc probe" ; timeout 300 python scripts/tc_probe.py > gpurun_out/tc_probe.log 2>&1 ; echo "tc_probe rc=$?" ; tail -12 gpurun_out/tc_probe.log
echo "=== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -40 gpurun_out/pytest_gpu.log
echo "=== bench 10M simt" ; timeout 1200 python bench.py --steps 3 --warmup 3 --gemm simt > gpurun_out/bench_10m_simt.log 2>&1 ; echo "bench10m rc=$?" ; tail -2 gpurun_out/bench_10m_simt.log
