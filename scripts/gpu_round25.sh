#!/bin/bash
# r1f, call 1: full GPU test suite (incl. the new CUDA-graph and narrow-hint tests), smoke, the
# headline bench with the in-process narrow-hint A/B, and graph-vs-eager on the small config shapes.
set -u
mkdir -p gpurun_out
echo "=== pytest gpu" ; timeout 420 python -m pytest tests -q -m gpu --timeout 150 > gpurun_out/pytest_gpu_r25.log 2>&1 ; echo "pytest rc=$?" ; tail -5 gpurun_out/pytest_gpu_r25.log
echo "=== smoke" ; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r25.log 2>&1 ; echo "smoke rc=$?" ; tail -3 gpurun_out/smoke_r25.log
echo "=== bench 10M default + narrow-hint A/B" ; timeout 330 python bench.py --steps 6 --no-cpu-baseline --no-e2e --ab-narrow-hint > gpurun_out/bench_10m_r25.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_10m_r25.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['peak_mem_gb'], d['loss']); print(d['roofline']); print(d['kernel_ms_per_step']); print('AB', d['narrow_hint_ab']); print('north', d['north_star_order']['ms_per_step'], d['north_star_order']['roofline']['frac'])"
run() { name=$1; shift; timeout 150 python bench.py --no-cpu-baseline --no-e2e --steps 20 "$@" > gpurun_out/bench_$name.log 2>&1; echo "$name rc=$?"; tail -1 gpurun_out/bench_$name.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  ms', round(d['ms_per_step'],4), 'eager ms', d.get('eager_ms_per_step'), 'edges/s', round(d['value']/1e6,1), 'M launches', d['gpu_launches'], 'loss', d['loss'])"; }
run cfg1_cora_shape_graph --graph --nodes 2708 --edges 10556 --fin 1433 --hidden 64 --nclass 7
run cfg2_squirrel_shape_graph --graph --nodes 5201 --edges 396846 --fin 2089 --hidden 64 --nclass 5 --model-type acmgcnp --structure-info 1
run cfg3_twitch_shape_graph --graph --nodes 168114 --edges 13595114 --fin 7 --hidden 256 --nclass 2 --model-type acmgcnp --flavour geometric --variant 1
run cfg4_arxiv_shape_graph --graph --nodes 169343 --edges 2315598 --fin 128 --hidden 256 --nclass 5 --model-type acmgcnpp --flavour geometric --variant 1
ls -la gpurun_out | head -30
