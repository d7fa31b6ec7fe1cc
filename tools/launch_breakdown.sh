#!/bin/bash
# Kernel-time breakdown of one E2EVN forward (B = 8) from an ncu launch list; output: gpurun_out/vn_launches.txt
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/vn_launches.csv python tools/prof_ops.py 8 2 vn > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/vn_launches.csv 2>&1 | tee gpurun_out/vn_launches.txt | head -40
