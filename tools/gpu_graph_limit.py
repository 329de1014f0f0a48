"""Graph replay vs direct launches on the bench frame (pool 32 Mi): RTB_GRAPH_MAX_POOL decides.  Run under gpurun."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracer_server_b200 as R
g = R.Scene.from_toml(os.path.join(ROOT, "tests/golden/scenes/flying_unicorn.toml"))
for lim in ("16777216", "1073741824", "16777216", "1073741824"):
    os.environ["RTB_GRAPH_MAX_POOL"] = lim
    g.render(1920, 1080, 256, seed=2)
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter(); g.render(1920, 1080, 256, seed=1); best = min(best, time.perf_counter() - t0)
    st = g.stats()
    print(f"graph limit {int(lim) >> 20} Mi: wall {best*1e3:.1f} ms iters {st['iterations']} -> {st['samples']/best/1e6:.1f} Msamples/s", flush=True)
