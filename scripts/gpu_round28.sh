#!/bin/bash
# r1f, call 2: tests, the default bench (both arms), launch list and ncu --set full of the row kernels
set -u
mkdir -p gpurun_out
echo "=== pytest gpu" ; timeout 420 python -m pytest tests -q -m gpu --timeout 150 > gpurun_out/pytest_gpu_r28.log 2>&1 ; echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_gpu_r28.log
echo "=== bench reference arm" ; timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_reference_r28.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_reference_r28.log | cut -c1-300
echo "=== bench default" ; timeout 420 python bench.py > gpurun_out/bench_r28.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_r28.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); n=d['north_star_order']
print({k:d[k] for k in ('value','ms_per_step','n_gpus','steps','warmup','dtype','gpu_launches','peak_mem_gb','loss','clocks')})
print('roofline', d['roofline']); print('north', n['ms_per_step'], n['roofline']['frac']); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step']); print('cpu', d['cpu_baseline']); print(d['kernel_ms_per_step'])"
echo "=== ncu launch list (timed region)" ; timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r28.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_r28.log 2>&1 ; echo "ncu list rc=$?"
echo "=== ncu full (row kernels, one step)" ; timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'spmm_|mix_bwd' -c 6 -o /tmp/prof_r28 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_r28.log 2>&1 ; echo "ncu full rc=$?"
ncu -i /tmp/prof_r28.ncu-rep --page raw --csv > gpurun_out/r28_raw.csv 2>/dev/null
ncu -i /tmp/prof_r28.ncu-rep --page details > gpurun_out/r28_details.txt 2>/dev/null
ncu -i /tmp/prof_r28.ncu-rep --page source --csv --kernel-name regex:mix_bwd > gpurun_out/r28_mix_bwd_source.csv 2>/dev/null
ncu -i /tmp/prof_r28.ncu-rep --page source --csv --kernel-name regex:spmm_mix_fwd > gpurun_out/r28_mix_fwd_source.csv 2>/dev/null
du -sh gpurun_out; ls -la gpurun_out | tail -15
