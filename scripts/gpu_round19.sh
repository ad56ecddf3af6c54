#!/bin/bash
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523"
echo "=== pytest gpu all" ; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_gpu.log
for R in 1 0; do
echo "=== bench 10M x1 mix_bwd ring=$R" ; ACMB200_MIXBWD_RING=$R timeout 900 python bench.py --steps 8 --no-cpu-baseline --no-e2e > gpurun_out/bench_10m_ring$R.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_10m_ring$R.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); n=d['north_star_order']; print(d['ms_per_step'], d['value'], 'north', n['ms_per_step'], n['roofline']['frac'], d['kernel_ms_per_step'])"
done
echo "=== bench 10M x$N" ; timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_10m_g${N}_r19.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_10m_g${N}_r19.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'ms', d['ms_per_step'], 'value', d['value'], 'north', (d['north_star_order'] or {}).get('ms_per_step'), d['kernel_ms_per_step'])"
