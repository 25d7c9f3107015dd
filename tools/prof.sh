#!/bin/bash
# ncu captures of the pair kernels on the configs[2] job (run under gpurun; outputs in gpurun_out/)
# usage: tools/prof.sh <tag> [atoms]
tag=${1:-x}; atoms=${2:-100000}
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${tag}.csv \
    python tools/profile_step.py $atoms 4 > gpurun_out/launch_run_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_search|k_classify|k_hscan|k_grid' -s 4 -c 4 \
    -o gpurun_out/prof_${tag} -f python tools/profile_step.py $atoms 4 > gpurun_out/prof_run_${tag}.log 2>&1
tail -2 gpurun_out/prof_run_${tag}.log
