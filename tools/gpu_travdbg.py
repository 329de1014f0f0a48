import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracer_server_b200 as R
os.environ["RTB_DEBUG_TRAVERSE"] = "1"
g = R.Scene.from_toml(os.path.join(ROOT, "tests/golden/scenes/flying_unicorn.toml"))
for steps in (4, 8, 16):
    for refill in (16, 22, 30):
        print("steps", steps, "refill", refill, flush=True)
        g.render(1920, 1080, 16, seed=1, count_work=True, tune_steps=steps, tune_refill=refill)
        st = g.stats()
        bvh = st["rays_bvh"] + st["shadow_bvh"]
        print(f"   nodes/bvh-ray {st['bvh_node_visits']/bvh:.1f} tris/bvh-ray {st['bvh_tri_tests']/bvh:.1f}", flush=True)
