#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest gpu all" ; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_gpu.log
for O in 2 3; do
echo "=== bench 10M mix_bwd occupancy=$O" ; ACMB200_MIXBWD_OCC=$O timeout 900 python bench.py --steps 8 --no-cpu-baseline --no-e2e > gpurun_out/bench_10m_occ$O.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_10m_occ$O.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); n=d['north_star_order']; print(d['ms_per_step'], d['value'], 'north', n['ms_per_step'], n['roofline']['frac'], d['kernel_ms_per_step'])"
done
echo "=== ncu full north-star order, async gather" ; timeout 1200 ncu --profile-from-start off --set full --clock-control none -k regex:'spmm_mix_fwd_kernel' -c 2 -o /tmp/prof_r18_ns python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --reorder off > gpurun_out/ncu_full_ns.log 2>&1 ; echo "ncu rc=$?"
ncu -i /tmp/prof_r18_ns.ncu-rep --page raw --csv > gpurun_out/r18_northstar_async_raw.csv 2>/dev/null
ncu -i /tmp/prof_r18_ns.ncu-rep --page details --kernel-name regex:spmm_mix_fwd > gpurun_out/r18_fused_async_details.txt 2>/dev/null
