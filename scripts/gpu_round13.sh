#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest gpu all" ; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_gpu.log
echo "=== ncu launch list (timed region, fwd+bwd)" ; timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r13.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list.log 2>&1 ; echo "ncu list rc=$?"
echo "=== ncu full (our kernels, timed region)" ; timeout 1800 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'spmm_|mix_bwd|tn_kernel|nt_kernel|cast_pad|nll_kernel' -c 24 -o gpurun_out/prof_r13 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1 ; echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log | cut -c1-200
echo "=== ncu full north-star order (fused SpMM+mix gather kernel)" ; timeout 1800 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'spmm_mix_fwd_kernel|spmm_t_kernel' -c 4 -o gpurun_out/prof_r13_northstar python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --reorder off > gpurun_out/ncu_full_ns.log 2>&1 ; echo "ncu full ns rc=$?"
