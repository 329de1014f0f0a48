"""usage: traffic_from_ncu.py <ncu --csv log> <stdout log with TRAFFIC_STATS> -> profiles/traffic.json fields (bytes per unit)."""
import csv, json, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
ci = {n: i for i, n in enumerate(hdr)}
tot = collections.Counter()
for r in rows[1:]:
    try:
        name, metric, unit, val = r[ci["Kernel Name"]], r[ci["Metric Name"]], r[ci["Metric Unit"]], float(r[ci["Metric Value"]].replace(",", ""))
    except Exception:
        continue
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    key = "k_shade" if "k_shade" in name else ("k_traverse" if "k_traverse" in name else ("k_generate" if "k_generate" in name else "other"))
    tot[key] += val * mult
st = None
for l in open(sys.argv[2], errors="replace"):
    if l.startswith("TRAFFIC_STATS "):
        st = json.loads(l[len("TRAFFIC_STATS "):])
vertices = st["rays_primary"] + st["rays_extension"]
out = {"k_shade_dram_bytes_per_vertex": tot["k_shade"] / vertices,
       "k_traverse_dram_bytes_per_bvh_ray": tot["k_traverse"] / (st["rays_bvh"] + st["shadow_bvh"]),
       "k_generate_dram_bytes_per_sample": tot["k_generate"] / st["samples"],
       "frame": st, "dram_bytes": dict(tot),
       "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over every launch of one 1920x1080 frame (tools/gpu_traffic.py), summed per kernel and divided by the frame's unit counts"}
print(json.dumps(out, indent=1))
