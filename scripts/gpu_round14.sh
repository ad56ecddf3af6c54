#!/bin/bash
# ncu evidence for the final kernels; reports are exported to CSV/text on the box (the .ncu-rep
# files are too large to bring back: gpurun_out is capped at 64 MiB).
set -u
mkdir -p gpurun_out
echo "=== ncu launch list (timed region, fwd+bwd)" ; timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r14.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list.log 2>&1 ; echo "ncu list rc=$?"
echo "=== ncu full (our kernels, timed region)" ; timeout 1800 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'spmm_|mix_bwd|tn_kernel|nt_kernel|cast_pad|nll_kernel' -c 24 -o /tmp/prof_r14 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1 ; echo "ncu full rc=$?"
ncu -i /tmp/prof_r14.ncu-rep --page raw --csv > gpurun_out/r14_default_raw.csv 2>/dev/null
ncu -i /tmp/prof_r14.ncu-rep --page details --kernel-name regex:spmm_agg_first > gpurun_out/r14_agg_first_details.txt 2>/dev/null
ncu -i /tmp/prof_r14.ncu-rep --page details --kernel-name regex:mix_bwd > gpurun_out/r14_mix_bwd_details.txt 2>/dev/null
ncu -i /tmp/prof_r14.ncu-rep --page details --kernel-name regex:tn_kernel > gpurun_out/r14_tn_gemm_details.txt 2>/dev/null
ncu -i /tmp/prof_r14.ncu-rep --page details --kernel-name regex:nt_kernel > gpurun_out/r14_nt_gemm_details.txt 2>/dev/null
ncu -i /tmp/prof_r14.ncu-rep --page source --csv --kernel-name regex:spmm_agg_first > gpurun_out/r14_agg_first_source.csv 2>/dev/null
echo "=== ncu full north-star order (fused SpMM+mix gather kernel)" ; timeout 1800 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'spmm_mix_fwd_kernel|spmm_t_kernel' -c 4 -o /tmp/prof_r14_ns python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --reorder off > gpurun_out/ncu_full_ns.log 2>&1 ; echo "ncu full ns rc=$?"
ncu -i /tmp/prof_r14_ns.ncu-rep --page raw --csv > gpurun_out/r14_northstar_raw.csv 2>/dev/null
ncu -i /tmp/prof_r14_ns.ncu-rep --page details --kernel-name regex:spmm_mix_fwd > gpurun_out/r14_fused_details.txt 2>/dev/null
ncu -i /tmp/prof_r14_ns.ncu-rep --page source --csv --kernel-name regex:spmm_mix_fwd > gpurun_out/r14_fused_source.csv 2>/dev/null
du -sh gpurun_out; ls -la gpurun_out
