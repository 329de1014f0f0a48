"""Smallest end-to-end case for compute-sanitizer (memcheck): every kernel, both estimators, probes, a job."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import raytracer_server_b200 as R
g = R.Scene.from_toml(os.path.join(ROOT, "tests/golden/scenes/flying_unicorn.toml"))
f = g.render(64, 48, 8, seed=1, pool_paths=2048)          # tiny pool: many regeneration rounds, graph mode
f2 = g.render(64, 48, 8, seed=1, use_mis=True, pool_paths=2048)
os.environ["RTB_NO_GRAPH"] = "1"
f3 = g.render(70, 50, 8, seed=1, rank=1, world=3, pool_paths=4096)
L = g.sample_radiance(64, 48, 8, np.arange(50) % 64, np.arange(50) % 48, np.arange(50) % 8, seed=2)
t = g.trace_primary(64, 48)
job = R.RenderJob(g, 64, 48, 8, passes=2, seed=3)
n = sum(1 for _ in job.messages()); job.close()
print("ok", f.mean(), f2.mean(), f3.mean(), L.mean(), (t["obj"] >= 0).mean(), n)
