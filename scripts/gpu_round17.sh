#!/bin/bash
# final 8-GPU validation: parity check + scaling bench (push exchange, aggregate-first, bf16 activations)
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
echo "=== dist_check x$N" ; timeout 600 $TR scripts/dist_check.py > gpurun_out/dist_check_$N.log 2>&1 ; echo "rc=$?" ; grep -E "dist_check|DIST_CHECK|Error|error|symmetric" gpurun_out/dist_check_$N.log | tail -10
echo "=== bench 10M x$N" ; timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_10m_g$N.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_10m_g$N.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'ms', d['ms_per_step'], 'value', d['value'], 'roof', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'north', (d['north_star_order'] or {}).get('ms_per_step'), 'mem', d['peak_mem_gb'], d['kernel_ms_per_step'])"
echo "=== bench 10M x1 (same box)" ; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_10m_g1_samebox.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_10m_g1_samebox.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); n=d['north_star_order']; print('N 1 ms', d['ms_per_step'], 'value', d['value'], 'roof', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'north', n['ms_per_step'], n['roofline']['frac'], 'cpu', d['cpu_baseline']['value'])"
