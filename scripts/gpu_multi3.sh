#!/bin/bash
# usage: gpurun --gpus N -- bash scripts/gpu_multi3.sh N   (r1f: multi-GPU parity + bench after the mix_bwd ring template)
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
echo "=== dist_check (fused peer push, default)" ; timeout 300 $TR scripts/dist_check.py > gpurun_out/dist_check_r1f_$N.log 2>&1 ; echo "rc=$?" ; grep -E "DIST_CHECK|Error|error|FAIL" gpurun_out/dist_check_r1f_$N.log | tail -6
SUM='
import json,sys
d=json.loads(sys.stdin.read()); print("N", d["n_gpus"], "ms", d["ms_per_step"], "value", d["value"], "roof", d["roofline"]["frac"], "exch", d["config"]["exchange"][-40:], d["kernel_ms_per_step"])'
echo "=== bench 10M x$N" ; timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_10m_r1f_g$N.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_10m_r1f_g$N.log | python -c "$SUM"
