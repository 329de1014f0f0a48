"""Extracts region statistics from the images the REFERENCE itself produced, so that the oracle (and the GPU
path) can be pinned to reference output on machines where /root/reference does not exist.

The reference holds no golden vectors and renders unseeded, but it ships three images that its own
renderer wrote:

  examples/cornell_box.png   600x450, written by the removed `--image ... --spp 64` CLI (render_examples.sh:8)
  examples/cubes.png         600x450, same CLI.  Its noise level, firefly count and clamp loss identify it as a
                             16-spp render (4 samples per sub-pixel): sample_pixel clamps every sub-pixel mean to
                             [0, 1] BEFORE averaging (src/server.rs:360), so the image mean depends on spp, and only
                             16 spp reproduces mean, noise and saturated-pixel count at once (see the table this
                             script prints; 64 spp is 2.8 % too bright, 1.6x too clean)
  raytracer.gif              screen recording of the browser client; the canvas sits 1:1 at (39, 23)-(639, 473) of
                             the last frame: cornell_box at "Samples Per Pixel = 4", after the recorder's 3-3-2 bit
                             palette (R, G: 8 levels, B: 4 levels, no dithering) and an unknown tone curve

What is stored (tests/golden/reference_pins.npz, ~40 KB; statistics, not the images):
  <name>_tiles   [9, 12, 3]   mean RGB of 50x50-pixel tiles
  <name>_blocks  [45, 60, 3]  mean RGB of 10x10-pixel blocks (float16)
  <name>_stats   [mean r, g, b, noise std (luma minus its 5x5 median), fireflies (luma > median + 80), pixels == 255]
  <name>_floor   oracle-vs-oracle noise floor of the same statistics (two seeds): [tile mean |rel|, tile max |rel|,
                 block rms rel] — what "equal up to Monte-Carlo noise" means at this spp

usage (needs /root/reference and the built oracle; ~4 CPU-minutes):  python tests/golden/make_reference_pins.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "reference_pins.npz")
W, H = 600, 450
GIF_CANVAS = (39, 23)   # top-left corner of the 600x450 canvas in raytracer.gif's frames


sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_metrics import blocks, compare_regions as compare, image_stats, tiles  # noqa: E402


if __name__ == "__main__":
    from PIL import Image

    from oracle import oracle as O

    threads = os.cpu_count() or 1
    out = {}
    cases = [("cornell_box", "examples/cornell_box.png", 64), ("cubes", "examples/cubes.png", 16)]
    for name, rel, spp in cases:
        img = np.asarray(Image.open(os.path.join(REF, rel)).convert("RGB")).astype(np.float64)
        assert img.shape == (H, W, 3)
        out[name + "_tiles"] = tiles(img)
        out[name + "_blocks"] = blocks(img).astype(np.float16)
        out[name + "_stats"] = image_stats(img)
        out[name + "_spp"] = spp
        sc = O.OracleScene.from_toml(os.path.join(ROOT, "tests", "golden", "scenes", name + ".toml"))
        sc.set_modes(O.ACCEL_OCTREE_FAITHFUL, O.EST_NEE)   # what the Rust binary runs (src/scene.rs:430-432, :217-229)
        a = sc.render(W, H, spp, seed=101, nthreads=-threads)["rgb8"].astype(np.float64)
        b = sc.render(W, H, spp, seed=102, nthreads=-threads)["rgb8"].astype(np.float64)
        out[name + "_floor"] = compare(a, b)
        print(f"{name}: reference image vs oracle at {spp} spp: {compare(a, img).round(4)}  oracle vs oracle: {out[name + '_floor'].round(4)}")
        print(f"   stats reference {out[name + '_stats'].round(2)}\n   stats oracle    {image_stats(a).round(2)}")
        if name == "cubes":   # the evidence for "16 spp"
            for s in (16, 32, 64):
                r = sc.render(W, H, s, seed=103, nthreads=-threads)["rgb8"].astype(np.float64)
                print(f"   cubes at {s} spp: vs reference {compare(r, img).round(4)}  stats {image_stats(r).round(2)}")
    g = Image.open(os.path.join(REF, "raytracer.gif"))
    g.seek(g.n_frames - 1)
    fr = np.asarray(g.convert("RGB")).astype(np.float64)
    x0, y0 = GIF_CANVAS
    can = fr[y0:y0 + H, x0:x0 + W]
    assert set(np.unique(can[..., 0])) <= {0., 36., 72., 108., 144., 180., 216., 252.} and set(np.unique(can[..., 2])) <= {0., 85., 170., 255.}
    out["gif4_tiles"] = tiles(can)
    out["gif4_spp"] = 4
    np.savez_compressed(OUT, **out)
    print(OUT, os.path.getsize(OUT), "bytes")
