#!/bin/bash
# usage: gpurun --gpus N -- bash scripts/gpu_multi.sh N
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
echo "=== dist_check" ; timeout 600 $TR scripts/dist_check.py > gpurun_out/dist_check_$N.log 2>&1 ; echo "rc=$?" ; grep -E "dist_check|DIST_CHECK|Error|error" gpurun_out/dist_check_$N.log | tail -8
echo "=== bench 10M x$N" ; timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_10m_g$N.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_10m_g$N.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'ms', d['ms_per_step'], 'value', d['value'], 'roof', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'north', (d['north_star_order'] or {}).get('ms_per_step'), 'mem', d['peak_mem_gb'], d['kernel_ms_per_step'])"
