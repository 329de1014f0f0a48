"""ACCEL_OCTREE_REFERENCE vs the LBVH: frame time on the bench frame, with and without ray binning.  Run under gpurun."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import raytracer_server_b200 as R
g = R.Scene.from_toml(os.path.join(ROOT, "tests/golden/scenes/flying_unicorn.toml"))
w, h, spp = 1920, 1080, 64
ref = None
for accel, bits, octm in ((0, 0, False), (1, 0, False), (1, 3, False), (1, 4, False), (1, 5, False), (1, 4, True), (1, 5, True)):
    g.render(w, h, 8, accel=accel, bin_bits=bits, bin_octant_major=octm)
    t0 = time.time(); f = g.render(w, h, spp, seed=1, accel=accel, bin_bits=bits, bin_octant_major=octm); dt = time.time() - t0
    st = g.stats()
    if accel == 1 and ref is None: ref = f
    d = int(np.abs(f.astype(int) - ref.astype(int)).max()) if accel == 1 else -1
    print(f"flying_unicorn {w}x{h}x{spp} accel {accel} bin bits {bits} octant-major {int(octm)}: dev {st['render_ms']:.1f} ms traverse {st['extend_ms']:.1f} bin {st['bin_ms']:.1f} shade {st['shade_ms']:.1f} "
          f"-> {st['samples']/dt/1e6:.1f} Msamples/s | max diff vs unbinned octree frame {d}", flush=True)
