#!/bin/bash
# One parametrised runner for every GPU-box call (replaces the per-call scripts of round 1):
#   gpurun [--gpus N] -- bash scripts/gpu_session.sh TAG STEP [STEP ...]
# Outputs go to gpurun_out/<TAG>_*.  Steps:
#   tests            pytest -m gpu (-x, as the driver runs it)
#   tests:<expr>     pytest -m gpu -k <expr>
#   qtests:<expr>    the same with a 300 s guard; aborts the rest of the session when it HANGS (new kernels)
#   smoke            __graft_entry__.build() + smoke()
#   ref              bench.py --impl reference (3 steps)
#   bench[:args]     bench.py [args]                      (single GPU; args comma separated, e.g. bench:--config,cfg3)
#   mbench:N[:args]  torchrun x N bench.py --gpus N [args]
#   dist:N           torchrun x N scripts/dist_check.py
#   nculist[:args]   ncu launch list (gpu__time_duration) of the timed region of bench.py [args]
#   ncufull:REGEX[:args]  ncu --set full of the kernels matching REGEX (one step), raw/details/source pages exported
set -u
TAG=$1; shift
mkdir -p gpurun_out
O=gpurun_out/$TAG
SUM='
import json,sys
try:
    d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1])
except Exception as e:
    print("no JSON line:", e); sys.exit(0)
print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","steps","warmup","dtype","gpu_launches","peak_mem_gb","loss","clocks")})
r=d.get("roofline") or {}
print("roofline", {k:r.get(k) for k in ("kernel","achieved","frac","avg_launch_ms","traffic","note")})
ns=d.get("north_star_order")
if ns: print("north-star order", ns["ms_per_step"], ns["roofline"]["frac"], ns.get("kernel_ms_per_step"))
e=d.get("e2e")
if e: print("e2e", e["value"], e["ms_per_step"])
print("cpu", (d.get("cpu_baseline") or {}).get("value"), "stock torch gpu", d.get("stock_torch_gpu"))
print("kernels", d.get("kernel_ms_per_step"))'
i=0
for STEP in "$@"; do
  i=$((i+1))
  KIND=${STEP%%:*}; REST=""; [[ "$STEP" == *:* ]] && REST=${STEP#*:}
  case $KIND in
    tests)
      if [ -n "$REST" ]; then timeout 1500 python -m pytest tests -x -q -m gpu -k "$REST" -s > ${O}_pytest_$i.log 2>&1
      else timeout 1700 python -m pytest tests -x -q -m gpu > ${O}_pytest_$i.log 2>&1; fi
      echo "=== [$STEP] rc=$?"; tail -5 ${O}_pytest_$i.log; grep -E "headline shape|ref_model_check" ${O}_pytest_$i.log | tail -20 ;;
    qtests)   # quick, guarded run of a few tests (new kernels): short timeout, abort the session on failure
      timeout 300 python -m pytest tests -x -q -m gpu -k "$REST" -s --timeout 120 > ${O}_qtests_$i.log 2>&1; RC=$?
      echo "=== [$STEP] rc=$RC"; tail -15 ${O}_qtests_$i.log | cut -c1-400
      if [ $RC -eq 124 ] || [ $RC -eq 137 ]; then echo "aborting session: guarded tests HUNG"; exit 0; fi ;;
    smoke)
      timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > ${O}_smoke.log 2>&1; echo "=== [smoke] rc=$?"; tail -3 ${O}_smoke.log ;;
    ref)
      timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 ${REST//,/ } > ${O}_bench_reference_$i.log 2>&1; echo "=== [$STEP] rc=$?"; tail -1 ${O}_bench_reference_$i.log | cut -c1-500 ;;
    ebench)   # ebench:VAR=VALUE[:args]  -- bench.py with one environment variable set (A/B of a knob on the same box)
      E=${REST%%:*}; A=""; [[ "$REST" == *:* ]] && A=${REST#*:}
      env $E timeout 1200 python bench.py ${A//,/ } > ${O}_bench_$i.log 2>&1; echo "=== [$STEP] rc=$?"; python -c "$SUM" < ${O}_bench_$i.log ;;
    bench)
      timeout 1200 python bench.py ${REST//,/ } > ${O}_bench_$i.log 2>&1; echo "=== [$STEP] rc=$?"; tail -3 ${O}_bench_$i.log | cut -c1-300 | grep -v '^{' ; python -c "$SUM" < ${O}_bench_$i.log ;;
    mbench)
      N=${REST%%:*}; A=""; [[ "$REST" == *:* ]] && A=${REST#*:}
      ACMB200_BENCH_WATCHDOG=${WATCHDOG:-240} timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+i)) bench.py --gpus $N ${A//,/ } > ${O}_bench_g${N}_$i.log 2>&1
      echo "=== [$STEP] rc=$?"; grep -iE "error|Traceback" ${O}_bench_g${N}_$i.log | head -5; python -c "$SUM" < ${O}_bench_g${N}_$i.log ;;
    dist)     # dist:N[:VAR=VALUE]  (e.g. dist:4:ACMB200_PUSH=0 for the NCCL exchange path)
      N=${REST%%:*}; E=""; [[ "$REST" == *:* ]] && E=${REST#*:}
      env $E timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+i)) scripts/dist_check.py > ${O}_dist_check_g${N}_$i.log 2>&1
      echo "=== [$STEP] rc=$?"; grep -E "DIST_CHECK|Error|FAIL" ${O}_dist_check_g${N}_$i.log | cut -c1-260 | tail -8; grep -c "> OK" ${O}_dist_check_g${N}_$i.log ;;
    nculist)
      timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches_$i.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-stock-torch ${REST//,/ } > ${O}_nculist_$i.log 2>&1
      echo "=== [$STEP] rc=$?"; wc -l ${O}_launches_$i.csv ;;
    ncufull)
      RX=${REST%%:*}; A=""; [[ "$REST" == *:* ]] && A=${REST#*:}
      timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"$RX" -c 20 -o /tmp/prof_${TAG}_$i python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-stock-torch ${A//,/ } > ${O}_ncufull_$i.log 2>&1
      echo "=== [$STEP] rc=$?"
      ncu -i /tmp/prof_${TAG}_$i.ncu-rep --page raw --csv > ${O}_ncu_raw_$i.csv 2>/dev/null
      ncu -i /tmp/prof_${TAG}_$i.ncu-rep --page details > ${O}_ncu_details_$i.txt 2>/dev/null
      ncu -i /tmp/prof_${TAG}_$i.ncu-rep --page source --csv > ${O}_ncu_source_$i.csv 2>/dev/null
      ls -la ${O}_ncu_* | tail -4 ;;
    *) echo "unknown step $STEP" ;;
  esac
done
du -sh gpurun_out
