#!/bin/bash
# Round profile pass (run under gpurun, 1 GPU).  Outputs land in gpurun_out/; tools/summarise_profiles.py turns
# them into the tracked profiles/ summaries.  Numbers printed under ncu are never bench values.
R=${1:-r02}
mkdir -p gpurun_out
# (1) launch list of the bench command itself (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 2200 --csv --log-file gpurun_out/${R}_bench_launches.csv \
    python bench.py --steps 1 --warmup 3 --batch 4 --no-cpu-baseline > gpurun_out/${R}_bench_under_ncu.log 2>&1
# (2) full capture of the hot kernels at the bench geometry (B=4, 15x320x320)
# one time step = row_dc320 (G8 output), conv5g, gru2, bh_fix_border, tc_kernel (conv3x3 d2), gru2, fin2: 7 launches
ncu --set full --clock-control none --import-source on -k regex:"tc_kernel|gru2_kernel|row_dc|conv5g|fin2|bh_fix" -s 7 -c 7 \
    -o gpurun_out/${R}_hot python tools/prof_ops.py 4 2 all > gpurun_out/${R}_hot.log 2>&1
ncu -i gpurun_out/${R}_hot.ncu-rep --page raw --csv > gpurun_out/${R}_hot_raw.csv 2>/dev/null
# (3) the real bench line, outside any profiler
python bench.py --steps 10 --warmup 3 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
tail -c 600 gpurun_out/${R}_bench.json
# (4) the IndRNN cell on BH activations (the cell base_cirim_run.yaml ships): one full capture + the tool's own timing
ncu --set full --clock-control none --import-source on -k regex:"ind2_kernel" -s 5 -c 1 -o gpurun_out/${R}_ind2 \
    python tools/time_tc2.py > gpurun_out/${R}_ind2.log 2>&1
ncu -i gpurun_out/${R}_ind2.ncu-rep --page raw --csv > gpurun_out/${R}_ind2_raw.csv 2>/dev/null
python tools/time_tc2.py > gpurun_out/${R}_time_tc2.log 2>&1
tail -6 gpurun_out/${R}_time_tc2.log
