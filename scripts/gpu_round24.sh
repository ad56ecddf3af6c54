#!/bin/bash
# gather mode 2 (cp.async.bulk ring): parity, then north-star order timing per gather mode
set -u
mkdir -p gpurun_out
echo "=== pytest gather modes" ; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "gather" > gpurun_out/pytest_gather.log 2>&1 ; echo "pytest rc=$?" ; tail -8 gpurun_out/pytest_gather.log
for G in 2 1; do
echo "=== bench 10M north-star order gather=$G" ; ACMB200_GATHER=$G timeout 600 python bench.py --steps 6 --reorder off --no-cpu-baseline --no-e2e > gpurun_out/bench_10m_ns_gather$G.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_10m_ns_gather$G.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline'], d.get('kernel_ms_per_step'))"
done
