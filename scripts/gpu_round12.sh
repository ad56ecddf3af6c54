#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest gpu all" ; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_gpu.log
echo "=== bench 10M default" ; timeout 900 python bench.py > gpurun_out/bench_10m_r12.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_10m_r12.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], (d['north_star_order'] or {}).get('ms_per_step'), d['kernel_ms_per_step'])"
echo "=== bench skewed 10M (skew 3)" ; timeout 900 python bench.py --skew 3 --no-cpu-baseline --no-e2e --steps 5 > gpurun_out/bench_10m_skew3.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_10m_skew3.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['nnz'], d['max_degree'], d['long_rows'], (d['north_star_order'] or {}).get('ms_per_step'), d['kernel_ms_per_step'])"
echo "=== bench cfg4-like (arXiv-year shape)" ; timeout 600 python bench.py --nodes 169343 --edges 2315598 --fin 128 --hidden 256 --nclass 5 --no-cpu-baseline > gpurun_out/bench_cfg4.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_cfg4.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['kernel_ms_per_step'])"
echo "=== ncu launch list (timed region)" ; timeout 1500 ncu --nvtx --nvtx-include "acm_timed_steps/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r12.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list.log 2>&1 ; echo "ncu list rc=$?"; tail -1 gpurun_out/ncu_list.log | cut -c1-200
echo "=== ncu full (our kernels, timed region)" ; timeout 1800 ncu --nvtx --nvtx-include "acm_timed_steps/" --set full --clock-control none --import-source on -k regex:'spmm_|mix_bwd|tn_kernel|nt_kernel|cast_pad|nll_kernel' -c 20 -o gpurun_out/prof_r12 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1 ; echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log | cut -c1-200
ls -la gpurun_out | tail -8
