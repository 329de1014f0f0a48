"""The C ABI from compiled code: examples/rtb_render_cli.c (C99, include/rtb200.h only) is built with the system C compiler
against librtb200.so.  Without a device it must fail loudly (no CPU fallback); on the GPU the PPM it assembles from the
streamed 60-pixel records is the frame the Python binding renders."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, SCENES, scene_path


@pytest.fixture(scope="module")
def cli(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("cli") / "rtb_render_cli")
    libdir = os.path.join(ROOT, "raytracer-server_b200")
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "rtb_render_cli.c"), "-L", libdir, "-lrtb200", f"-Wl,-rpath,{libdir}", "-o", exe], check=True)
    return exe


def test_c_consumer_builds_and_refuses_without_a_device(cli, tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([cli, scene_path("cornell_box"), os.path.join(SCENES, "assets"), "64", "48", "8", str(tmp_path / "o.ppm")],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr and not (tmp_path / "o.ppm").exists()
    assert subprocess.run([cli], capture_output=True).returncode == 2          # usage


@pytest.mark.gpu
@pytest.mark.parametrize("scene,accel", [("cubes", "lbvh"), ("flying_unicorn", "octree")])
def test_c_consumer_renders_the_same_frame(cli, rtb, gpu_scene, tmp_path, scene, accel):
    W, H, spp, seed = 200, 150, 16, 77
    out = tmp_path / "o.ppm"
    r = subprocess.run([cli, scene_path(scene), os.path.join(SCENES, "assets"), str(W), str(H), str(spp), str(out), accel, str(seed)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert f"{H * 4} records, {W * H} pixels, {W * H * spp} samples" in r.stdout          # windows of 60, 60, 60, 20
    raw = out.read_bytes()
    head = f"P6\n{W} {H}\n255\n".encode()
    assert raw.startswith(head)
    frame = np.frombuffer(raw[len(head):], dtype=np.uint8).reshape(H, W, 3).astype(int)
    want = gpu_scene(scene).render(W, H, spp, seed=seed, accel=rtb.ACCEL_OCTREE_REFERENCE if accel == "octree" else 0).astype(int)
    assert np.abs(frame - want).max() <= 1
