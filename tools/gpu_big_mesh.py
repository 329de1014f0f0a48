"""A mesh far larger than any reference asset: build time and traversal agreement of the hierarchy builders.  Run under gpurun."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import raytracer_server_b200 as R
n = int(os.environ.get("NTRI", "1000000"))
rng = np.random.default_rng(1)
# a bumpy sphere shell of small triangles + a few huge ones (mixed sizes are what binning has to cope with)
c = rng.normal(size=(n, 3)); c /= np.linalg.norm(c, axis=1, keepdims=True); c *= 30 + rng.normal(size=(n, 1))
tris = c[:, None, :] + rng.normal(size=(n, 3, 3)) * 0.08
nh = int(os.environ.get("NHUGE", "50"))
if nh: tris[:nh] = rng.uniform(-40, 40, size=(nh, 3, 3))
objs = [{"brdf": ("diffuse", (0.7, 0.7, 0.7)), "geometry": ("mesh", tris)},
        {"emitted": (20, 20, 20), "brdf": ("diffuse", (0, 0, 0)), "geometry": ("sphere", (0, 90, 0), 2.0)}]
m = int(os.environ.get("NRAYS", "2000000"))
org = rng.uniform(-50, 50, size=(m, 3)).astype(np.float32)
d = rng.normal(size=(m, 3)); d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
ref = None
for mode in os.environ.get("MODES", "lbvh,,ploc").split(","):
    if mode: os.environ["RTB_BVH"] = mode
    else: os.environ.pop("RTB_BVH", None)
    ts = []
    for rep in range(2):
        t0 = time.perf_counter(); g = R.Scene.from_objects((0, 0, 120), (0, 0, -1), objs); ts.append(time.perf_counter() - t0)
    g.trace_rays(org, d)
    tr = 1e9
    for rep in range(3):
        t0 = time.perf_counter(); r = g.trace_rays(org, d); tr = min(tr, time.perf_counter() - t0)
    os.environ["RTB_NO_GRAPH"] = "1"
    g.render(512, 512, 16, seed=1)
    g.render(512, 512, 64, seed=1); st = g.stats()
    if ref is None: ref = r
    same = bool(np.array_equal(ref["obj"], r["obj"]) and np.array_equal(ref["t"], r["t"]))
    print(f"{n} triangles RTB_BVH={mode or 'sah (default)'}: create {min(ts)*1e3:.0f} ms (build {g.info.build_ms:.1f} ms) nodes {g.info.bvh_nodes} depth {g.info.bvh_depth} | "
          f"{m} rays {tr*1e3:.1f} ms, hits {(r['tri'] >= 0).mean():.3f}, same as Karras: {same} | 512x512x64 frame: traverse {st['extend_ms']:.1f} ms of {st['render_ms']:.1f}", flush=True)
    del g
