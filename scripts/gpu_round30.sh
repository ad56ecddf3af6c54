#!/bin/bash
# r1f, final sanity of the committed tree: what the driver runs first (GPU tests with -x, smoke)
set -u
mkdir -p gpurun_out
echo "=== pytest gpu -x" ; timeout 400 python -m pytest tests -x -q -m gpu --timeout 200 > gpurun_out/pytest_gpu_r30.log 2>&1 ; echo "pytest rc=$?" ; tail -2 gpurun_out/pytest_gpu_r30.log
echo "=== smoke" ; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r30.log 2>&1 ; echo "smoke rc=$?" ; tail -3 gpurun_out/smoke_r30.log
