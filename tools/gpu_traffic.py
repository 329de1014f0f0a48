"""One flying_unicorn frame for the DRAM-traffic capture (run under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`):
prints the frame's unit counts, which tools/traffic_from_ncu.py divides the per-kernel byte sums by."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracer_server_b200 as R
g = R.Scene.from_toml(os.path.join(ROOT, "tests/golden/scenes/flying_unicorn.toml"))
w, h, spp = 1920, 1080, int(sys.argv[1]) if len(sys.argv) > 1 else 16
g.render(w, h, spp, seed=3)
st = g.stats()
print("TRAFFIC_STATS " + json.dumps({k: st[k] for k in ("samples", "rays_primary", "rays_extension", "rays_shadow", "rays_bvh", "shadow_bvh",
                                                          "paths_queued", "iterations")}), flush=True)
