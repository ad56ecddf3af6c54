#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest gpu all" ; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -4 gpurun_out/pytest_gpu.log
run() { name=$1; shift; timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 6 "$@" > gpurun_out/bench_$name.log 2>&1; echo "$name rc=$?"; tail -1 gpurun_out/bench_$name.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  ms', round(d['ms_per_step'],3), 'edges/s', round(d['value']/1e6,1), 'M', d['kernel_ms_per_step'])"; }
run cfg5_acmgcnp_geo_v1 --model-type acmgcnp --flavour geometric --variant 1
run cfg5_acmgcnp_geo_v0 --model-type acmgcnp --flavour geometric --variant 0
run cfg5_default
