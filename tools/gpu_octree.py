"""ACCEL_OCTREE_REFERENCE vs the LBVH: frame time on the bench frame.  Run under gpurun."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracer_server_b200 as R
g = R.Scene.from_toml(os.path.join(ROOT, "tests/golden/scenes/flying_unicorn.toml"))
for w, h, spp in ((600, 450, 64), (1920, 1080, 256)):
    for accel in (0, 1):
        g.render(w, h, 8, accel=accel)
        t0 = time.time(); g.render(w, h, spp, seed=1, accel=accel); dt = time.time() - t0
        st = g.stats()
        print(f"flying_unicorn {w}x{h}x{spp} accel {accel}: wall {dt*1e3:.1f} ms dev {st['render_ms']:.1f} ms traverse {st['extend_ms']:.1f} shade {st['shade_ms']:.1f} "
              f"-> {st['samples']/dt/1e6:.1f} Msamples/s; octree nodes {g.info.octree_nodes}", flush=True)
