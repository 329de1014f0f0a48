"""Finer sweep: traversal knobs around the optimum, pool size at the bench frame (256 spp).  Run under gpurun."""
import os, sys, time, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracer_server_b200 as R
g = R.Scene.from_toml(os.path.join(ROOT, "tests/golden/scenes/flying_unicorn.toml"))
w, h = 1920, 1080
g.render(w, h, 8)
def run(spp, **kw):
    g.render(w, h, spp, seed=1, **kw)
    st = g.stats()
    it = st["iterations"]
    return st["samples"]/st["render_ms"]/1e3, st["extend_ms"]/it*1e3, st["shade_ms"]/it*1e3, it, st["extend_ms"], st["shade_ms"], st["render_ms"]
for refill, steps in itertools.product((26, 28, 30), (6, 8, 10, 12)):
    r = run(32, tune_refill=refill, tune_steps=steps)
    print(f"refill {refill:2d} steps {steps:4d}: {r[0]:6.1f} Msamples/s  traverse {r[1]:5.0f} us/iter shade {r[2]:5.0f} us/iter iters {r[3]}", flush=True)
for P in (1 << 24, 1 << 25, 1 << 26):
    g.render(w, h, 8, pool_paths=P)
    r = run(256, pool_paths=P)
    print(f"256 spp pool {P>>20}M: {r[0]:6.1f} Msamples/s traverse {r[4]:6.1f} ms shade {r[5]:6.1f} ms total {r[6]:6.1f} ms iters {r[3]}", flush=True)
