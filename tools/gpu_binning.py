"""A/B of the coherence binning in front of k_traverse (rtb_params.tuning[3]).  Run under gpurun."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("RTB_DEBUG_TRAVERSE", "1")
import numpy as np
import raytracer_server_b200 as R
SC = os.path.join(ROOT, "tests/golden/scenes")
scenes = sys.argv[1:] or ["flying_unicorn"]
for name in scenes:
    g = R.Scene.from_toml(os.path.join(SC, name + ".toml"))
    w, h, spp = (1920, 1080, 256) if name == "flying_unicorn" else (600, 450, 256)
    g.render(w, h, 8)
    base = None
    for bits, octm in ((0, False), (3, False), (4, False), (5, False), (4, True), (5, True), (0, False)):
        best = None
        for rep in range(2):
            f = g.render(w, h, spp, seed=1, bin_bits=bits, bin_octant_major=octm)
            st = g.stats()
            if best is None or st["render_ms"] < best["render_ms"]:
                best = st
        if base is None:
            base = f
        d = np.abs(f.astype(int) - base.astype(int))
        it = best["iterations"]
        print(f"{name} {w}x{h}x{spp} bits {bits} octant_major {int(octm)}: dev {best['render_ms']:.1f} ms = {best['samples']/best['render_ms']/1e3:.1f} Msamples/s | "
              f"traverse {best['extend_ms']:.1f} bin {best['bin_ms']:.1f} shade {best['shade_ms']:.1f} | iters {it} | frame max diff vs off {d.max()}", flush=True)
    # SIMD-slot accounting of the counting build (stderr)
    for bits in (0, 4, 5):
        print(f"-- counting build, bits {bits}", flush=True)
        sys.stderr.flush()
        g.render(960, 540, 64, seed=1, bin_bits=bits, count_work=True)
        st = g.stats()
        print(f"   node visits {st['bvh_node_visits']} tri tests {st['bvh_tri_tests']}", flush=True)
