"""Node visits / triangle tests per LBVH ray (counting build of k_traverse) for each hierarchy builder.  Run under gpurun."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracer_server_b200 as R
os.environ["RTB_DEBUG_TRAVERSE"] = "1"
for mode in ("", "ploc", "lbvh"):
    if mode: os.environ["RTB_BVH"] = mode
    else: os.environ.pop("RTB_BVH", None)
    g = R.Scene.from_toml(os.path.join(ROOT, "tests/golden/scenes/flying_unicorn.toml"))
    g.render(1920, 1080, 16, seed=1, count_work=True)
    st = g.stats()
    rays = st["rays_bvh"] + st["shadow_bvh"]
    print(f"RTB_BVH={mode or 'sah (default)'}: {rays} LBVH rays, {st['bvh_node_visits']/rays:.2f} node visits and {st['bvh_tri_tests']/rays:.2f} triangle tests per ray "
          f"(nodes {g.info.bvh_nodes}, depth {g.info.bvh_depth})", flush=True)
