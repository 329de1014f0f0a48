"""A/B of the inline tail (RTB_INLINE_TAIL = paths left below which k_shade traverses the LBVH itself).  Run under gpurun."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import raytracer_server_b200 as R
SC = os.path.join(ROOT, "tests/golden/scenes")
for name, w, h, spp in (("flying_unicorn", 1920, 1080, 256), ("cubes", 600, 450, 256), ("flying_unicorn", 600, 450, 64), ("flying_unicorn", 3840, 2160, 32)):
    g = R.Scene.from_toml(os.path.join(SC, name + ".toml"))
    g.render(w, h, 8)
    ref = None
    for tail in [int(x) for x in os.environ.get("TAILS", "0,4096,32768,262144,1048576,0,32768").split(",")]:
        os.environ["RTB_INLINE_TAIL"] = str(tail)
        best = None
        for rep in range(3):
            t0 = time.perf_counter(); f = g.render(w, h, spp, seed=1); dt = time.perf_counter() - t0
            st = g.stats(); st["wall"] = dt
            if best is None or dt < best["wall"]: best = st
        if ref is None: ref = f
        d = int(np.abs(f.astype(int) - ref.astype(int)).max())
        print(f"{name} {w}x{h}x{spp} tail<{tail}: wall {best['wall']*1e3:.2f} ms dev {best['render_ms']:.2f} ms iters {best['iterations']} launches {best['kernel_launches']} "
              f"-> {best['samples']/best['wall']/1e6:.1f} Msamples/s | max diff {d} rays {best['rays_extension']+best['rays_shadow']}", flush=True)
