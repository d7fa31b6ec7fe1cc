#!/bin/bash
# Short GPU-box visit: tensor-core parity, kernel timings, the headline bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
echo "== tc tests"; timeout 900 python -m pytest tests/test_gpu_tc.py -x -q -s --timeout=300 > gpurun_out/pytest_tc.log 2>&1; echo "tc rc=$?"; grep -E "tc parity|passed|failed|Error|error" gpurun_out/pytest_tc.log | tail -30
echo "== misc tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "many_small or cirim_full or metrics" --timeout=300 > gpurun_out/pytest_misc.log 2>&1; echo "misc rc=$?"; tail -5 gpurun_out/pytest_misc.log
echo "== timings"; for b in 4 16; do timeout 300 python tools/time_tc2.py $b; done 2>&1 | tee gpurun_out/time_tc2.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "ms/ts", d["roofline"]["ms_per_time_step"], "dc", d["roofline_dc"]["frac"], "shares", d["kernel_shares"], "parity", d["cpu_baseline"].get("parity_rel_l2_vs_cuda"))
except Exception as e:
    print("bench parse failed", e)
PY
tail -5 gpurun_out/bench.err
