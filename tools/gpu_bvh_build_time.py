"""Build time of each hierarchy builder (second build of the same scene: no first-use costs).  Run under gpurun."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracer_server_b200 as R
p = os.path.join(ROOT, "tests/golden/scenes/flying_unicorn.toml")
g0 = R.Scene.from_toml(p)
blob_objects = None
for mode in ("", "ploc", "lbvh", "sah_host"):
    if mode: os.environ["RTB_BVH"] = mode
    else: os.environ.pop("RTB_BVH", None)
    ts = []
    for rep in range(4):
        t0 = time.perf_counter(); g = R.Scene.from_toml(p); ts.append(time.perf_counter() - t0)
    print(f"RTB_BVH={mode or 'sah (default)'}: scene load + upload + build {min(ts)*1e3:.1f} ms (nodes {g.info.bvh_nodes}, depth {g.info.bvh_depth})", flush=True)
