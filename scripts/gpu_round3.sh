#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -8 gpurun_out/pytest_gpu.log
echo "=== bench 10M tcgen05" ; timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_10m_tc.log 2>&1 ; echo "bench rc=$?" ; tail -1 gpurun_out/bench_10m_tc.log | cut -c1-3000
echo "=== ncu launch list" ; timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'acm|at::|elementwise|reduce|softmax|nll|index|vectorized' -c 1500 --csv --log-file gpurun_out/launches_10m.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list.log 2>&1 ; echo "ncu list rc=$?"; tail -2 gpurun_out/ncu_list.log | cut -c1-300
echo "=== ncu full" ; timeout 1800 ncu --set full --clock-control none --import-source on -k regex:'spmm_mix_fwd_kernel|spmm_t_kernel|mix_bwd_kernel|tn_kernel|nt_kernel|cast_pad' -s 11 -c 11 -o gpurun_out/prof_r1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1 ; echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out
