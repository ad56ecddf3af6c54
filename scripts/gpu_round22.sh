#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest gpu all (pack/unpack kernels)" ; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -5 gpurun_out/pytest_gpu.log
run() { name=$1; shift; timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 10 "$@" > gpurun_out/bench_$name.log 2>&1; echo "$name rc=$?"; tail -1 gpurun_out/bench_$name.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  ms', round(d['ms_per_step'],3), 'edges/s', round(d['value']/1e6,1), 'M launches', d['gpu_launches'], d['kernel_ms_per_step'])"; }
run cfg1_cora_shape --nodes 2708 --edges 10556 --fin 1433 --hidden 64 --nclass 7
run cfg2_squirrel_shape --nodes 5201 --edges 396846 --fin 2089 --hidden 64 --nclass 5 --model-type acmgcnp --structure-info 1
echo "=== ncu LN forward kernel" ; timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'spmm_mix_fwd_kernel' -c 1 -o /tmp/prof_ln python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --model-type acmgcnp --flavour geometric --variant 1 > gpurun_out/ncu_ln.log 2>&1 ; echo "ncu rc=$?"
ncu -i /tmp/prof_ln.ncu-rep --page details > gpurun_out/r22_fused_ln_details.txt 2>/dev/null
ncu -i /tmp/prof_ln.ncu-rep --page source --csv > gpurun_out/r22_fused_ln_source.csv 2>/dev/null
echo "=== ncu non-LN variant-1 forward kernel" ; timeout 1200 ncu --profile-from-start off --set full --clock-control none -k regex:'spmm_mix_fwd_kernel' -c 1 -o /tmp/prof_v1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --variant 1 > gpurun_out/ncu_v1.log 2>&1 ; echo "ncu rc=$?"
ncu -i /tmp/prof_v1.ncu-rep --page details > gpurun_out/r22_fused_v1_details.txt 2>/dev/null
