#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest gpu all" ; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -6 gpurun_out/pytest_gpu.log
echo "=== smoke" ; timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -3 gpurun_out/smoke.log
echo "=== bench 10M default" ; timeout 900 python bench.py > gpurun_out/bench_10m_r10.log 2>&1 ; echo "rc=$?" ; tail -1 gpurun_out/bench_10m_r10.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], (d['north_star_order'] or {}).get('ms_per_step'), d['kernel_ms_per_step'])"
echo "=== bench twitch-like (N=168114, E=13.6M, Fin=7, hidden=256, 2 classes)" ; timeout 600 python bench.py --nodes 168114 --edges 13595114 --fin 7 --hidden 256 --nclass 2 --no-cpu-baseline > gpurun_out/bench_cfg3.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_cfg3.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['kernel_ms_per_step'])"
