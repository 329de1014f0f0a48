#!/bin/bash
# usage (here): tools/ncu_read.sh <tag> <mangled-kernel-substring> [top]   -- stall summary + hot source lines of gpurun_out/prof_<tag>.ncu-rep
tag=$1; kern=$2; top=${3:-30}
cd "$(dirname "$0")/../gpurun_out"
ncu -i prof_$tag.ncu-rep --page raw --csv > /tmp/${tag}_raw.csv 2>/dev/null
ncu -i prof_$tag.ncu-rep --page source --csv > /tmp/${tag}_src.csv 2>/dev/null
python - <<PY
import csv
rows=list(csv.reader(open('/tmp/${tag}_raw.csv')))
hdr=rows[0]; r=rows[2]
want=['gpu__time_duration.sum','launch__registers_per_thread','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','sm__warps_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active']
for k,v in zip(hdr,r):
    try: f=float(v.replace(',',''))
    except: continue
    if k in want or ('average_warps_issue_stalled' in k and f>0.05): print(k,v)
PY
mkdir -p /tmp/cubin && cd /tmp/cubin && cuobjdump -xelf all /root/repo/raytracer-server_b200/librtb200.so > /dev/null && nvdisasm -g -c engine.sm_100a.cubin > /tmp/engine.asm
python /root/repo/tools/ncu_hot_lines.py /tmp/${tag}_src.csv $kern /tmp/engine.asm $top
