"""BASELINE config 2 with the dead "MIS" estimator: cubes 600x450 256 spp (and unicorn 1080p 32 spp).  Run under gpurun."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracer_server_b200 as R
for name, w, h, spp in (("cubes", 600, 450, 256), ("flying_unicorn", 1920, 1080, 32)):
    g = R.Scene.from_toml(os.path.join(ROOT, "tests/golden/scenes", name + ".toml"))
    g.render(w, h, 8, use_mis=True)
    best = 1e9
    for i in range(3):
        t0 = time.time(); g.render(w, h, spp, seed=i, use_mis=True); best = min(best, time.time() - t0)
    st = g.stats()
    print(f"{name} {w}x{h}x{spp} MIS-dead: {st['samples']/best/1e6:.1f} Msamples/s wall, dev {st['render_ms']:.1f} ms traverse {st['extend_ms']:.1f} shade {st['shade_ms']:.1f} iters {st['iterations']}", flush=True)
