#!/bin/bash
# final single-GPU validation of the round: what the driver runs (tests, smoke, both bench arms)
set -u
mkdir -p gpurun_out
echo "=== pytest gpu (-x as the driver does)" ; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_final.log 2>&1 ; echo "pytest rc=$?" ; tail -4 gpurun_out/pytest_gpu_final.log
echo "=== smoke" ; timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke_final.log 2>&1 ; echo "smoke rc=$?" ; tail -3 gpurun_out/smoke_final.log
echo "=== bench reference arm" ; timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_reference_final.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_reference_final.log | cut -c1-400
echo "=== bench default" ; timeout 900 python bench.py > gpurun_out/bench_final.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_final.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); n=d['north_star_order']
print({k:d[k] for k in ('value','ms_per_step','n_gpus','steps','warmup','dtype','gpu_launches','peak_mem_gb','loss','clocks')})
print('roofline', d['roofline']); print('north', n['ms_per_step'], n['roofline']['frac'], n['roofline']['traffic']); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step']); print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])"
