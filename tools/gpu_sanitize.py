"""Smallest end-to-end case for compute-sanitizer (memcheck / racecheck): every kernel, every estimator and accel mode,
probes, pixel lists, binning, a banded streaming job and a progressive one."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import raytracer_server_b200 as R
from raytracer_server_b200.host import sample_pixels
g = R.Scene.from_toml(os.path.join(ROOT, "tests/golden/scenes/flying_unicorn.toml"))
f = g.render(64, 48, 8, seed=1, pool_paths=2048)          # tiny pool: many regeneration rounds, graph mode
f2 = g.render(64, 48, 8, seed=1, use_mis=True, pool_paths=2048)
f4 = g.render(64, 48, 8, seed=1, estimator=R.EST_MIS_BALANCE, pool_paths=2048)
f5 = g.render(64, 48, 8, seed=1, accel=R.ACCEL_OCTREE_REFERENCE, pool_paths=2048)
f6 = g.render(64, 48, 8, seed=1, bin_bits=4, pool_paths=2048)
os.environ["RTB_NO_GRAPH"] = "1"
f3 = g.render(70, 50, 8, seed=1, rank=1, world=3, pool_paths=4096)
f7 = g.render(64, 48, 8, seed=1, bin_bits=5, bin_octant_major=True, pool_paths=4096)
f8 = g.render(64, 48, 8, seed=1, accel=R.ACCEL_OCTREE_REFERENCE, use_mis=True, pool_paths=4096)
L = g.sample_radiance(64, 48, 8, np.arange(50) % 64, np.arange(50) % 48, np.arange(50) % 8, seed=2)
L2 = g.sample_radiance(64, 48, 8, np.arange(50) % 64, np.arange(50) % 48, np.arange(50) % 8, seed=2, accel=R.ACCEL_OCTREE_REFERENCE)
P = sample_pixels(np.arange(40) % 64, np.arange(40) % 48, 64, 48, 8, g, seed=2)
t = g.trace_primary(64, 48)
t2 = g.trace_rays(np.zeros((10, 3), np.float32) + [50, 40, 200], np.tile([0.0, -0.2, -1.0], (10, 1)).astype(np.float32), accel=R.ACCEL_OCTREE_REFERENCE)
del os.environ["RTB_NO_GRAPH"]
os.environ["RTB_BAND_TILE_ROWS"] = "1"
job = R.RenderJob(g, 100, 70, 8, seed=3)                  # 3 bands, 2 workers
n1 = sum(1 for _ in job.messages()); job.close()
job = R.RenderJob(g, 64, 48, 8, passes=2, seed=3)
n = sum(1 for _ in job.messages()); job.close()
job = R.RenderJob(g, 200, 150, 64, seed=4)                # cancelled mid-way
job.stop(); job.close()
print("ok", f.mean(), f2.mean(), f3.mean(), f4.mean(), f5.mean(), f6.mean(), f7.mean(), f8.mean(), L.mean(), L2.mean(), P.mean(), (t["obj"] >= 0).mean(), t2["obj"][:2], n1, n)
