#!/bin/bash
# usage (under gpurun): tools/ncu_one.sh <tag> <kernel-regex> <launch-skip> [scene]
# one `ncu --set full` capture of one launch of a kernel inside tools/gpu_perf.py; report -> gpurun_out/prof_<tag>.ncu-rep
tag=$1; kern=$2; skip=$3; scene=${4:-flying_unicorn}
RTB_PERF_NOWARM=1 ncu --set full --import-source on --clock-control none -k regex:$kern -s $skip -c 1 -f \
    -o gpurun_out/prof_$tag python tools/gpu_perf.py $scene > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log
