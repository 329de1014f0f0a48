"""BASELINE.json's image gate at full spp: "converged 4096-spp images must match within 1 % mean relative error
per channel (and PSNR >= 40 dB), with MIS on and off".

Two families of committed oracle frames (600x450 = the reference server's size; generator: tests/golden/make_converged.py):

* `<scene>_{nee,mis}_..._4096spp`: f64, EXACT nearest hit (the reference's brute-force branch, src/geometry.rs:887-903).
  The GPU renders the same seed under the shared RNG contract and must reproduce the frame.
* `<scene>_octree_nee_..._1024spp`: f64, the reference's REAL mesh path — the early-exit octree it always builds
  (src/scene.rs:430-432, src/geometry.rs:1263-1273).  The distance of the GPU frame (exact nearest hit, same
  seed, same spp) to these is the distance to what the Rust binary renders: the "reference-defect floor".  On cubes the
  octree never returns a non-nearest triangle (the oracle's two modes agree bit for bit), on flying_unicorn it does.
Every number is recorded through parity_log (profiles/parity_rNN.json)."""
import glob
import os

import numpy as np
import pytest

from conftest import ROOT
from parity_metrics import channel_mre, pixel_mre, psnr

pytestmark = pytest.mark.gpu
ALL = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "converged", "*.npz")))
EXACT = [f for f in ALL if "_octree_" not in os.path.basename(f)]
OCTREE = [f for f in ALL if "_octree_" in os.path.basename(f)]


def scene_of(path):
    return os.path.basename(path).split("_octree_")[0].split("_nee_")[0].split("_mis_")[0]


def render_like(gpu_scene, path):
    z = np.load(path)
    w, h, spp, seed, est = int(z["width"]), int(z["height"]), int(z["spp"]), int(z["seed"]), int(z["estimator"])
    got = gpu_scene(scene_of(path)).render(w, h, spp, seed=seed, use_mis=bool(est))
    return got, z["rgb8"], spp, est


def measure(got, gold):
    d = np.abs(got.astype(int) - gold.astype(int))
    return {"pixel_mre": pixel_mre(got, gold, blur=1), "pixel_mre_blur5": pixel_mre(got, gold), "channel_mre": channel_mre(got, gold),
            "psnr": psnr(got, gold), "frac_beyond_3_levels": (d > 3).mean(), "max_level_diff": int(d.max())}


@pytest.mark.skipif(not EXACT, reason="no converged golden frames committed yet")
@pytest.mark.parametrize("path", EXACT, ids=[os.path.basename(f)[:-4] for f in EXACT])
def test_converged_frame_matches_oracle(gpu_scene, parity_log, path):
    got, gold, spp, est = render_like(gpu_scene, path)
    m = measure(got, gold)
    parity_log(f"gpu/converged_exact/{os.path.basename(path)[:-4]}", spp=spp, **m)
    assert (m["pixel_mre"] < 0.01).all(), f"per-pixel mean relative error per channel {m['pixel_mre']}"
    assert (m["channel_mre"] < 0.01).all()
    assert m["psnr"] >= 40.0
    assert m["frac_beyond_3_levels"] < 0.01


@pytest.mark.skipif(not OCTREE, reason="no octree-faithful golden frames committed yet")
@pytest.mark.parametrize("path", OCTREE, ids=[os.path.basename(f)[:-4] for f in OCTREE])
def test_distance_to_the_reference_octree(gpu_scene, parity_log, path):
    # GPU = exact nearest hit; golden = the reference's early-exit octree; same seed, same spp, same estimator
    got, gold, spp, est = render_like(gpu_scene, path)
    m = measure(got, gold)
    d = np.abs(got.astype(int) - gold.astype(int)).max(axis=2)
    ys, xs = np.nonzero(d > 3)
    bbox = [int(xs.min()), int(ys.min()), int(xs.max()), int(ys.max())] if xs.size else None
    parity_log(f"gpu/reference_defect_floor/{os.path.basename(path)[:-4]}", spp=spp, pixels_beyond_3_levels=int(xs.size),
               bbox_of_those_pixels=bbox, **m)
    if scene_of(path) == "cubes":     # the octree is exact on the two cubes: the strict same-seed gate applies
        assert m["max_level_diff"] <= 3 and m["psnr"] >= 50.0
        assert (m["pixel_mre_blur5"] < 0.01).all() and (m["channel_mre"] < 0.01).all()
    else:
        # flying_unicorn: the early-exit octree returns non-nearest triangles and the Rust binary renders a
        # visibly darker mesh.  Measured (profiles/parity_r2.json): channel means 2.0-2.3 % apart, per-pixel MRE 4-6.5 %
        # (2.6-3.6 % after a 5x5 blur), PSNR 30.3 dB, 10 % of the pixels off by more than 3 levels, all of it on the mesh,
        # its shadow and its mirror image.  No exact nearest-hit renderer can be closer than this to the binary; the bounds
        # below only guard the measurement against regressions.
        assert (m["channel_mre"] < 0.03).all() and (m["channel_mre"] > 0.01).all()
        assert 28.0 <= m["psnr"] <= 33.0
        ys, xs = np.nonzero(d > 40)              # the gross differences sit on the mesh (screen centre-left), not on the walls
        assert 150 < np.median(xs) < 350 and 100 < np.median(ys) < 400


@pytest.mark.skipif(not OCTREE, reason="no octree-faithful golden frames committed yet")
@pytest.mark.parametrize("path", OCTREE, ids=[os.path.basename(f)[:-4] for f in OCTREE])
def test_octree_mode_reproduces_the_reference_frame(rtb, gpu_scene, parity_log, path):
    # ACCEL_OCTREE_REFERENCE: the GPU walks the reference's own octrees — BASELINE.json's image gate against what the Rust
    # binary renders, flying_unicorn included
    z = np.load(path)
    w, h, spp, seed = int(z["width"]), int(z["height"]), int(z["spp"]), int(z["seed"])
    got = gpu_scene(scene_of(path)).render(w, h, spp, seed=seed, accel=rtb.ACCEL_OCTREE_REFERENCE)
    m = measure(got, z["rgb8"])
    parity_log(f"gpu/octree_mode_frame/{os.path.basename(path)[:-4]}", spp=spp, **m)
    assert (m["pixel_mre"] < 0.01).all() and (m["channel_mre"] < 0.01).all()
    assert m["psnr"] >= 40.0 and m["frac_beyond_3_levels"] < 0.01
