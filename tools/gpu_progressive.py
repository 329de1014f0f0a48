"""BASELINE config 5: progressive cornell_box 600x450, 1 sample per pixel per frame; frames/s and time to converge."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import raytracer_server_b200 as R

def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)

g = R.Scene.from_toml(os.path.join(ROOT, "tests/golden/scenes/cornell_box.toml"))
W, H = 600, 450
ref = g.render(W, H, 4096, seed=99)
g.render(W, H, 8)
for spp in (256, 1024):
    job = R.RenderJob(g, W, H, spp, passes=spp, seed=1)   # one (k, sub-pixel) per pass = 1 sample / pixel / frame
    t0 = time.time()
    marks = {}
    n = 0
    for i, f in job.frames():
        n += 1
        if (i + 1) % 4 == 0 and (i + 1) in (4, 16, 64, 256, 1024):
            p = psnr(f, ref)
            marks[i + 1] = (time.time() - t0, p)
    dt = time.time() - t0
    job.close()
    print(f"progressive cornell_box {W}x{H}: {n} frames of 1 sample/pixel in {dt*1e3:.0f} ms -> {n/dt:.0f} frames/s; "
          + "; ".join(f"{k} spp: {v[0]*1e3:.0f} ms, PSNR vs 4096 spp {v[1]:.1f} dB" for k, v in marks.items()), flush=True)
