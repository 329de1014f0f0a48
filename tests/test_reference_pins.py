"""Pins against images the REFERENCE ITSELF produced (tests/golden/reference_pins.npz, made by
tests/golden/make_reference_pins.py from /root/reference/examples/*.png and raytracer.gif).

The reference renders unseeded, so equality means "statistically the same image": region means, noise level,
firefly count and clamp loss of a fresh render must sit on the oracle-vs-oracle noise floor stored in the fixture.
  * CPU tests pin the ORACLE (oracle/rt_oracle.cpp) in the reference's real mode: octree traversal + live NEE.
  * GPU tests pin the PRODUCT directly: librtb200 renders the same scenes at the same spp with its own seeds.
The per-sub-pixel clamp of sample_pixel (src/server.rs:360) makes the image mean depend on spp, which is why a
16-spp render is needed for cubes.png (see the generator's docstring) — and why these pins are sharp: camera,
tent filter, estimator, clamp order and gamma all have to be right for the numbers to agree at the 0.3 % level.
"""
import os

import numpy as np
import pytest

from conftest import NCPU, ROOT, scene_path
from parity_metrics import compare_regions, image_stats, tiles, tone_curve_residual

PINS = np.load(os.path.join(ROOT, "tests", "golden", "reference_pins.npz"))
W, H = 600, 450
CASES = [("cornell_box", 64), ("cubes", 16)]


def check_against_reference(img, name, who, parity_log):
    """img: a fresh RGB8 render of `name` at the pinned spp.  Gates are multiples of the stored noise floor."""
    img = img.astype(np.float64)
    floor = PINS[name + "_floor"]                 # oracle vs oracle: [tile mean |rel|, tile max |rel|, block rms rel]
    t_ref, b_ref = PINS[name + "_tiles"], PINS[name + "_blocks"].astype(np.float64)
    t = tiles(img)
    rel = np.abs(t - t_ref) / t_ref
    from parity_metrics import blocks

    brel = (blocks(img) - b_ref) / np.maximum(b_ref, 8.0)
    block_rms = float(np.sqrt((brel ** 2).mean()))
    st, st_ref = image_stats(img), PINS[name + "_stats"]
    parity_log(f"{who}/reference_image/{name}", spp=int(PINS[name + "_spp"]), tile_mean_rel=rel.mean(), tile_max_rel=rel.max(),
               block_rms_rel=block_rms, noise_floor=floor, channel_means=st[:3], channel_means_reference=st_ref[:3],
               noise_std=st[3], noise_std_reference=st_ref[3], fireflies=st[4], fireflies_reference=st_ref[4],
               saturated=st[5], saturated_reference=st_ref[5])
    assert rel.mean() < 1.35 * floor[0] + 5e-4, f"tile means: {rel.mean():.4f} vs noise floor {floor[0]:.4f}"
    assert rel.max() < max(2.2 * floor[1], 0.03)
    assert block_rms < 1.12 * floor[2], f"10x10 block rms {block_rms:.4f} vs noise floor {floor[2]:.4f}"
    assert np.abs(st[:3] - st_ref[:3]).max() / st_ref[:3].min() < 0.004        # channel means within 0.4 %
    assert abs(st[3] - st_ref[3]) / st_ref[3] < 0.03                             # same noise level => same spp, same estimator
    assert abs(st[5] - st_ref[5]) / st_ref[5] < 0.08                             # saturated pixels: light disc + clamp loss
    assert 0.4 * st_ref[4] - 10 < st[4] < 2.0 * st_ref[4] + 25                   # fireflies (a Poisson count)


def check_against_gif(img, who, parity_log):
    rms, mx = tone_curve_residual(tiles(img.astype(np.float64)), PINS["gif4_tiles"])
    parity_log(f"{who}/reference_gif/cornell_box", spp=4, tone_fit_rms_rel=rms, tone_fit_max_rel=mx)
    assert rms < 0.02 and mx < 0.06


# ---------------------------------------------------------------- the oracle (CPU)
@pytest.mark.parametrize("name,spp", CASES)
def test_oracle_reproduces_reference_image(oracle_mod, parity_log, name, spp):
    assert int(PINS[name + "_spp"]) == spp
    sc = oracle_mod.OracleScene.from_toml(scene_path(name))
    sc.set_modes(oracle_mod.ACCEL_OCTREE_FAITHFUL, oracle_mod.EST_NEE)   # the Rust binary: src/scene.rs:430-432, :217-229
    img = sc.render(W, H, spp, seed=2024, nthreads=-NCPU)["rgb8"]
    check_against_reference(img, name, "oracle", parity_log)


def test_oracle_reproduces_reference_gif_structure(oracle_mod, parity_log):
    # raytracer.gif, last frame: cornell_box at 4 spp through the recorder's 3-3-2 palette and tone curve
    sc = oracle_mod.OracleScene.from_toml(scene_path("cornell_box"))
    sc.set_modes(oracle_mod.ACCEL_OCTREE_FAITHFUL, oracle_mod.EST_NEE)
    check_against_gif(sc.render(W, H, 4, seed=7, nthreads=-NCPU)["rgb8"], "oracle", parity_log)
    # the comparison has teeth: another scene does not fit
    other = oracle_mod.OracleScene.from_toml(scene_path("cubes"))
    rms, mx = tone_curve_residual(tiles(other.render(W, H, 4, seed=7, nthreads=-NCPU)["rgb8"].astype(np.float64)), PINS["gif4_tiles"])
    assert rms > 0.04 and mx > 0.2


def test_wrong_spp_is_detected():
    # the pins distinguish 16 from 64 spp on cubes (what identified examples/cubes.png): use the stored statistics of
    # the OTHER image as a stand-in for "a render with the wrong clamp loss"
    a, b = PINS["cornell_box_stats"], PINS["cubes_stats"]
    assert abs(a[3] - b[3]) / b[3] > 0.2


# ---------------------------------------------------------------- the product (GPU)
@pytest.mark.gpu
@pytest.mark.parametrize("name,spp", CASES)
def test_gpu_reproduces_reference_image(gpu_scene, parity_log, name, spp):
    img = gpu_scene(name).render(W, H, spp, seed=77)
    check_against_reference(img, name, "gpu", parity_log)


@pytest.mark.gpu
def test_gpu_reproduces_reference_gif_structure(gpu_scene, parity_log):
    check_against_gif(gpu_scene("cornell_box").render(W, H, 4, seed=78), "gpu", parity_log)
