#!/bin/bash
set -u
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 10 "$@" > gpurun_out/bench_$name.log 2>&1; echo "$name rc=$?"; tail -1 gpurun_out/bench_$name.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  ms', round(d['ms_per_step'],3), 'edges/s', round(d['value']/1e6,1), 'M  nnz', d['nnz'], d['kernel_ms_per_step'])"; }
run cfg3_acmgcnp_geo_v1 --nodes 168114 --edges 13595114 --fin 7 --hidden 256 --nclass 2 --model-type acmgcnp --flavour geometric --variant 1
run cfg3_acmgcnp_geo_v1_s1 --nodes 168114 --edges 13595114 --fin 7 --hidden 256 --nclass 2 --model-type acmgcnp --flavour geometric --variant 1 --structure-info 1
run cfg3_acmgcnp_geo_v0 --nodes 168114 --edges 13595114 --fin 7 --hidden 256 --nclass 2 --model-type acmgcnp --flavour geometric --variant 0
run cfg4_acmgcnpp_geo_v1 --nodes 169343 --edges 2315598 --fin 128 --hidden 256 --nclass 5 --model-type acmgcnpp --flavour geometric --variant 1
run cfg2_squirrel_shape --nodes 5201 --edges 396846 --fin 2089 --hidden 64 --nclass 5 --model-type acmgcnp --structure-info 1
run cfg1_cora_shape --nodes 2708 --edges 10556 --fin 1433 --hidden 64 --nclass 7
run cfg5_variant1 --variant 1 --steps 5
run cfg5_acmgcnp_geo_v1 --model-type acmgcnp --flavour geometric --variant 1 --steps 5
