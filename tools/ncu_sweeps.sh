#!/bin/bash
# ncu captures of the fused sweep kernels (development tool; run under gpurun)
# usage: tools/ncu_sweeps.sh [tag] [sweep_one.py tuning args...]
set -x
TAG=${1:-sweeps}; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sweep_ -c 24 --csv --log-file gpurun_out/launches_$TAG.csv python tools/sweep_one.py 2048 1024 6 "$@" > gpurun_out/ncu_${TAG}1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep_ -s 8 -c 2 -o gpurun_out/prof_$TAG -f python tools/sweep_one.py 2048 1024 6 "$@" > gpurun_out/ncu_${TAG}2.log 2>&1
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv --print-source sass > gpurun_out/prof_${TAG}_sass.csv 2>/dev/null
ncu -i gpurun_out/prof_$TAG.ncu-rep --page details > gpurun_out/prof_${TAG}_details.txt 2>/dev/null
tail -n 3 gpurun_out/ncu_${TAG}1.log gpurun_out/ncu_${TAG}2.log
