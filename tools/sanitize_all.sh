#!/bin/bash
# compute-sanitizer passes over the hot path (CIRIM BH engine incl. the bulk-copy first conv / tap-GEMM final conv / G8 DC
# output, E2EVN incl. the tensor-core U-Net convs).  Logs: gpurun_out/sanitizer/*.log (copied to profiles/<round>_sanitizer/).
mkdir -p gpurun_out/sanitizer
S=/usr/local/cuda/bin/compute-sanitizer
run() { # name tool args...
  local name=$1 tool=$2; shift 2
  timeout 900 $S --tool $tool --log-file gpurun_out/sanitizer/${name}.raw "$@" > gpurun_out/sanitizer/${name}.out 2>&1
  echo "rc=$?" >> gpurun_out/sanitizer/${name}.out
  grep -E "COMPUTE-SANITIZER|ERROR SUMMARY|RACECHECK SUMMARY|Error|error:" gpurun_out/sanitizer/${name}.raw | head -20 > gpurun_out/sanitizer/${name}.log
  tail -2 gpurun_out/sanitizer/${name}.out >> gpurun_out/sanitizer/${name}.log
  echo "== $name"; cat gpurun_out/sanitizer/${name}.log
}
ONLY=${1:-}   # optional substring filter on the run name (e.g. `indrnn`)
_run() { run "$@"; }
run_if() { case "$1" in *"$ONLY"*) _run "$@";; esac; }
run_if memcheck_indrnn_b2 memcheck python tools/sanitize_run.py 2 320 indrnn
run_if synccheck_indrnn_b1 synccheck python tools/sanitize_run.py 1 64 indrnn
run_if racecheck_indrnn_b1_64x320 racecheck python tools/sanitize_run.py 1 64 indrnn
[ -n "$ONLY" ] && exit 0
run memcheck_cirim_b2 memcheck python tools/sanitize_run.py 2 320 cirim
run memcheck_e2evn_b2 memcheck python tools/sanitize_run.py 2 320 vn
# initcheck is 100x+ slower than the others: run it alone with a longer limit if needed
# run initcheck_cirim_b1 initcheck python tools/sanitize_run.py 1 64 cirim
run synccheck_cirim_b1 synccheck python tools/sanitize_run.py 1 64 cirim
run synccheck_e2evn_b1 synccheck python tools/sanitize_run.py 1 64 vn
run racecheck_cirim_b1_64x320 racecheck python tools/sanitize_run.py 1 64 cirim
run racecheck_e2evn_b1_64x320 racecheck python tools/sanitize_run.py 1 64 vn
