"""Sweep of the traversal knobs / pool size on flying_unicorn.  Run under gpurun."""
import os, sys, time, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracer_server_b200 as R
g = R.Scene.from_toml(os.path.join(ROOT, "tests/golden/scenes/flying_unicorn.toml"))
w, h, spp = 1920, 1080, 32
g.render(w, h, 8)
def run(**kw):
    g.render(w, h, spp, seed=1, **kw)
    st = g.stats()
    it = st["iterations"]
    return st["samples"]/st["render_ms"]/1e3, st["extend_ms"]/it*1e3, st["shade_ms"]/it*1e3, it
for refill, steps in itertools.product((16, 22, 28, 32), (2, 4, 8, 16, 1000)):
    r = run(tune_refill=refill, tune_steps=steps)
    print(f"refill {refill:2d} steps {steps:4d}: {r[0]:6.1f} Msamples/s  traverse {r[1]:5.0f} us/iter shade {r[2]:5.0f} us/iter iters {r[3]}", flush=True)
for P in (1 << 21, 1 << 22, 1 << 23, 1 << 24):
    g.render(w, h, 8, pool_paths=P)
    r = run(pool_paths=P)
    print(f"pool {P>>20}M: {r[0]:6.1f} Msamples/s traverse {r[1]:5.0f} shade {r[2]:5.0f} us/iter iters {r[3]}", flush=True)
