"""Pool size on the small configs (C1 / C2 / 600x450 flying_unicorn): default (samples / 8) vs larger pools.  Run under gpurun."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracer_server_b200 as R
SC = os.path.join(ROOT, "tests/golden/scenes")
for name, w, h, spp in (("cornell_box", 600, 450, 64), ("cubes", 600, 450, 256), ("flying_unicorn", 600, 450, 64), ("flying_unicorn", 600, 450, 256), ("cornell_box", 600, 450, 4)):
    g = R.Scene.from_toml(os.path.join(SC, name + ".toml"))
    for P in (0, 1 << 20, 1 << 21, 1 << 22, 1 << 23, 1 << 24):
        g.render(w, h, spp, seed=2, pool_paths=P)
        best = None
        for rep in range(3):
            t0 = time.perf_counter(); g.render(w, h, spp, seed=1, pool_paths=P); dt = time.perf_counter() - t0
            st = g.stats(); st["wall"] = dt
            if best is None or dt < best["wall"]: best = st
        print(f"{name} {w}x{h}x{spp} pool {P >> 20 if P else 'default'}Mi: wall {best['wall']*1e3:.2f} ms dev {best['render_ms']:.2f} iters {best['iterations']} -> {best['samples']/best['wall']/1e6:.1f} Msamples/s", flush=True)
