#!/bin/bash
# One GPU-box visit: TC parity first (fast fail), full parity suite, role profile, bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== tc tests"; timeout 900 python -m pytest tests/test_gpu_tc.py -x -q -s --timeout=600 > gpurun_out/pytest_tc.log 2>&1; echo "tc rc=$?"; grep -E "tc parity|passed|failed|Error|error" gpurun_out/pytest_tc.log | tail -40
echo "== roles"; timeout 300 python tools/tc_roles.py 4 > gpurun_out/roles.log 2>&1; echo "roles rc=$?"; tail -12 gpurun_out/roles.log
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
