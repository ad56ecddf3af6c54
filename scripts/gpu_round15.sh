#!/bin/bash
set -u
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
echo "=== single-GPU bench (north-star kernel regression check)" ; timeout 600 python bench.py --steps 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_10m_r15.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_10m_r15.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); n=d['north_star_order']; print(d['ms_per_step'], d['value'], d['roofline']['frac'], 'north', n['ms_per_step'], n['roofline']['avg_launch_ms'], n['roofline']['frac'], d['kernel_ms_per_step'])"
echo "=== dist_check push=1" ; timeout 600 $TR scripts/dist_check.py > gpurun_out/dist_check_push_$N.log 2>&1 ; echo "rc=$?" ; grep -E "dist_check|DIST_CHECK|Error|error|symmetric" gpurun_out/dist_check_push_$N.log | tail -12
echo "=== dist_check push=0" ; ACMB200_PUSH=0 timeout 600 $TR scripts/dist_check.py > gpurun_out/dist_check_nopush_$N.log 2>&1 ; echo "rc=$?" ; grep -E "DIST_CHECK|Error|error" gpurun_out/dist_check_nopush_$N.log | tail -3
for P in 1 0; do
echo "=== bench 10M x$N push=$P" ; ACMB200_PUSH=$P timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_10m_g${N}_push$P.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_10m_g${N}_push$P.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'ms', d['ms_per_step'], 'value', d['value'], 'north', (d['north_star_order'] or {}).get('ms_per_step'), d['kernel_ms_per_step'])"
done
