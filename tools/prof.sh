#!/bin/bash
# ncu captures of the step's kernels on the configs[2] job (run under gpurun; outputs in gpurun_out/)
# usage: tools/prof.sh <tag> [atoms]
#   launches_<tag>.csv      every launch of 4 steps with its device time (cold cache, serialised)
#   prof_<tag>.ncu-rep      --set full, caches flushed before every replay (stalls, issue, source lines)
#   traffic_<tag>.csv       DRAM and L2 bytes of the same kernels with --cache-control none: the lists one kernel hands to
#                           the next stay in L2 as in a real step, so DRAM bytes here are what the step really moves
tag=${1:-x}; atoms=${2:-100000}
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${tag}.csv \
    python tools/profile_step.py $atoms 4 > gpurun_out/launch_run_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_search|k_classify|k_hscan|k_grid' -s 4 -c 4 \
    -o gpurun_out/prof_${tag} -f python tools/profile_step.py $atoms 4 > gpurun_out/prof_run_${tag}.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none --cache-control none \
    -k regex:'k_search|k_classify|k_hscan|k_grid' -s 8 -c 8 --csv --log-file gpurun_out/traffic_${tag}.csv \
    python tools/profile_step.py $atoms 5 > gpurun_out/traffic_run_${tag}.log 2>&1
tail -2 gpurun_out/prof_run_${tag}.log
