#!/bin/bash
# ncu evidence for the round (run under gpurun, one GPU): launch list of the bench command and one
# --set full capture of the sweep kernels.  Outputs under gpurun_out/; summarise with tools/profile_summarise.py.
TAG=${1:-r1i}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extra --no-api-loop > gpurun_out/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep_ -s 8 -c 4 -o gpurun_out/prof_$TAG -f \
    python tools/sweep_one.py 2048 1024 6 > gpurun_out/ncu_full_$TAG.log 2>&1
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
tail -n 2 gpurun_out/ncu_launches_$TAG.log gpurun_out/ncu_full_$TAG.log
