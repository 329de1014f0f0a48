import os, sys, time, itertools
sys.path.insert(0, "/root/repo")
import raytracer_server_b200 as R
g = R.Scene.from_toml("/root/repo/tests/golden/scenes/flying_unicorn.toml")
w, h, spp = 1920, 1080, 32
g.render(w, h, 8, accel=1, bin_bits=0)
for refill, steps in itertools.product((16, 24, 28, 31), (2, 4, 8, 16, 32)):
    best = 1e9
    for rep in range(2):
        t0 = time.perf_counter(); g.render(w, h, spp, seed=1, accel=1, bin_bits=0, tune_refill=refill, tune_steps=steps); best = min(best, time.perf_counter() - t0)
    print(f"octree refill {refill} steps {steps}: {best*1e3:.1f} ms", flush=True)
