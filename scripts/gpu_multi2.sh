#!/bin/bash
# usage: gpurun --gpus N -- bash scripts/gpu_multi2.sh N
# multi-GPU parity (incl. hidden-256 wide pushes) + bench with the NVSwitch-multicast push on and off
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
echo "=== dist_check (multicast on)" ; ACMB200_MULTICAST=1 timeout 600 $TR scripts/dist_check.py > gpurun_out/dist_check_mc_$N.log 2>&1 ; echo "rc=$?" ; grep -E "dist_check|DIST_CHECK|Error|error" gpurun_out/dist_check_mc_$N.log | tail -14
echo "=== dist_check (multicast off, default)" ; timeout 600 $TR scripts/dist_check.py > gpurun_out/dist_check_uc_$N.log 2>&1 ; echo "rc=$?" ; grep -E "DIST_CHECK|Error|error" gpurun_out/dist_check_uc_$N.log | tail -4
SUM='
import json,sys
d=json.loads(sys.stdin.read()); print("N", d["n_gpus"], "ms", d["ms_per_step"], "value", d["value"], "roof", d["roofline"]["frac"], "exch", d["config"]["exchange"][-40:], "north", (d.get("north_star_order") or {}).get("ms_per_step"), d["kernel_ms_per_step"])'
echo "=== bench 10M x$N multicast on" ; ACMB200_MULTICAST=1 timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_10m_mc_g$N.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_10m_mc_g$N.log | python -c "$SUM"
echo "=== bench 10M x$N multicast off (default)" ; timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_10m_uc_g$N.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_10m_uc_g$N.log | python -c "$SUM"
