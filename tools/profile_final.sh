#!/bin/bash
# usage (under gpurun): tools/profile_final.sh <tag>      -- everything profiles/ holds for one build
# 1. bench.py, both arms (never under a profiler)                          -> gpurun_out/<tag>_bench.json, <tag>_bench_reference.json
# 2. launch list of a bench run (gpu__time_duration, cold, serialised)     -> gpurun_out/<tag>_launches.csv
# 3. DRAM bytes + warp instructions of EVERY launch of one 1920x1080x16 frame -> gpurun_out/<tag>_traffic.csv / .log (tools/traffic_from_ncu.py)
# 4. one `ncu --set full` capture of k_traverse + k_shade (iteration 12 of a 1920x1080x32 flying_unicorn frame) -> gpurun_out/prof_<tag>.ncu-rep
# 2-4 launch the kernels directly (RTB_NO_GRAPH=1), like the bench frame's 32 Mi-slot pool does
tag=$1
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || exit 1
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err
RTB_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --spp 32 > gpurun_out/${tag}_ncu_launches.log 2>&1
RTB_NO_GRAPH=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_traffic.csv python tools/gpu_traffic.py 16 > gpurun_out/${tag}_traffic.log 2>&1
RTB_NO_GRAPH=1 RTB_PERF_NOWARM=1 ncu --set full --import-source on --clock-control none -k 'regex:k_shade|k_traverse' -s 24 -c 2 -f \
    -o gpurun_out/prof_${tag} python tools/gpu_perf.py flying_unicorn > gpurun_out/${tag}_ncu_full.log 2>&1
tail -c 600 gpurun_out/${tag}_bench.json
