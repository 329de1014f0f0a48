"""usage: traffic_from_ncu.py <ncu --csv log of tools/gpu_traffic.py> <its stdout log with TRAFFIC_STATS> [<k_shade --set full raw csv> <k_traverse raw csv>] [commit]
-> profiles/traffic.json: DRAM bytes and warp instructions per unit of the two hot kernels (ncu over EVERY launch of one
1920x1080 frame, summed per kernel, divided by the frame's counted units), plus issue / pipe / lane figures of one launch each."""
import csv, json, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
ci = {n: i for i, n in enumerate(hdr)}
dram, inst, launches = collections.Counter(), collections.Counter(), collections.Counter()
for r in rows[1:]:
    try:
        name, metric, unit, val = r[ci["Kernel Name"]], r[ci["Metric Name"]], r[ci["Metric Unit"]], float(r[ci["Metric Value"]].replace(",", ""))
    except Exception:
        continue
    key = "k_shade" if "k_shade" in name else ("k_traverse" if "k_traverse" in name else ("k_generate" if "k_generate" in name else "other"))
    if metric.startswith("dram__bytes"):
        dram[key] += val * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    elif metric == "smsp__inst_executed.sum":
        inst[key] += val
        launches[key] += 1
st = None
for l in open(sys.argv[2], errors="replace"):
    if l.startswith("TRAFFIC_STATS "):
        st = json.loads(l[len("TRAFFIC_STATS "):])
vertices = st["rays_primary"] + st["rays_extension"]
bvh_rays = st["rays_bvh"] + st["shadow_bvh"]
out = {"k_shade_dram_bytes_per_vertex": dram["k_shade"] / vertices,
       "k_traverse_dram_bytes_per_bvh_ray": dram["k_traverse"] / bvh_rays,
       "k_generate_dram_bytes_per_sample": dram["k_generate"] / st["samples"],
       "k_shade_warp_inst_per_vertex": inst["k_shade"] / vertices,
       "k_traverse_warp_inst_per_bvh_ray": inst["k_traverse"] / bvh_rays,
       "k_generate_warp_inst_per_sample": inst["k_generate"] / st["samples"],
       "frame": st, "dram_bytes": dict(dram), "warp_instructions": dict(inst), "launches": dict(launches),
       "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum over every launch of one 1920x1080 frame "
              "(tools/gpu_traffic.py), summed per kernel and divided by the frame's unit counts"}
if len(sys.argv) > 4:
    want = {"smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct", "smsp__thread_inst_executed_per_inst_executed.ratio": "lanes_per_instruction",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
            "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct", "launch__registers_per_thread": "registers",
            "l1tex__t_sector_hit_rate.pct": "l1_hit_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak"}
    for kern, path in (("k_shade", sys.argv[3]), ("k_traverse", sys.argv[4])):
        rr = list(csv.reader(open(path)))
        h, v = rr[0], rr[2]
        for k, x in zip(h, v):
            if k in want:
                out[f"{kern}_{want[k]}"] = float(x.replace(",", ""))
    out["issue_source"] = "ncu --set full of one launch each (iteration 12 of a 1920x1080x32 frame): profiles/*_ncu_full_raw.csv of the same build"
if len(sys.argv) > 5:
    out["commit"] = sys.argv[5]
print(json.dumps(out, indent=1))
