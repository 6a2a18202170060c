#!/bin/bash
# ncu captures of the fused sweep kernels (development tool; run under gpurun)
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sweep_ -c 24 --csv --log-file gpurun_out/launches_sweeps.csv python tools/sweep_one.py 2048 1024 6 > gpurun_out/ncu_sweeps1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep_ -s 8 -c 4 -o gpurun_out/prof_sweeps -f python tools/sweep_one.py 2048 1024 6 > gpurun_out/ncu_sweeps2.log 2>&1
ncu -i gpurun_out/prof_sweeps.ncu-rep --page raw --csv > gpurun_out/prof_sweeps_raw.csv 2>/dev/null
tail -3 gpurun_out/ncu_sweeps1.log gpurun_out/ncu_sweeps2.log
