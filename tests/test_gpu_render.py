"""Behaviour of the render entry points on the GPU: frame conventions, size-independent properties
at BASELINE.json sizes, tile sharding, streaming jobs, cancellation, error paths."""
import ctypes as C
import threading
import time

import numpy as np
import pytest

from conftest import scene_path

pytestmark = pytest.mark.gpu


def test_spp_rule_and_black_frame(gpu_scene):
    g = gpu_scene("cornell_box")
    f = g.render(64, 48, 3)                     # spp / 4 == 0 -> black (src/server.rs:332)
    assert (f == 0).all() and g.stats()["samples"] == 0
    f7 = g.render(64, 48, 7, seed=1)
    assert g.stats()["samples"] == 64 * 48 * 4  # effective spp = 4 * (spp / 4)
    assert np.array_equal(f7, g.render(64, 48, 4, seed=1))


def test_frame_orientation_and_light(gpu_scene):
    # row 0 is the TOP of the screen (message row y, sampler row height - y - 1, src/server.rs:177-181):
    # the light hangs near the ceiling, the floor is at the bottom
    f = gpu_scene("cornell_box").render(200, 150, 16, seed=2).astype(int)
    ys = np.argwhere(f.min(axis=2) >= 250)[:, 0]
    assert ys.size > 10 and ys.mean() < 75
    left, right = f[40:110, 5:25].reshape(-1, 3).mean(0), f[40:110, 175:195].reshape(-1, 3).mean(0)
    assert left[0] > left[2] * 1.3 and right[2] > right[0] * 1.3      # red wall left, blue wall right


def test_counted_rays_per_sample(gpu_scene):
    g = gpu_scene("cubes")
    g.render(600, 450, 16, seed=3)
    st = g.stats()
    rays = st["rays_primary"] + st["rays_extension"] + st["rays_shadow"]
    assert st["samples"] == st["rays_primary"] == 600 * 450 * 16
    assert 24 < rays / st["samples"] < 34      # ~30 trace_ray calls per sample (zero-throughput paths culled)
    assert st["kernel_launches"] >= 4 * st["iterations"] > 0


def test_seed_changes_noise_not_mean(gpu_scene):
    g = gpu_scene("cubes")
    a, b = g.render(300, 225, 64, seed=1).astype(int), g.render(300, 225, 64, seed=2).astype(int)
    assert not np.array_equal(a, b)
    assert np.allclose(a.reshape(-1, 3).mean(0), b.reshape(-1, 3).mean(0), rtol=0.004)
    c = g.render(300, 225, 64, seed=1).astype(int)
    assert np.abs(a - c).max() <= 1            # same seed: identical up to fp32 atomic ordering


@pytest.mark.parametrize("use_mis", [False, True])
def test_frame_independent_of_the_path_pool(gpu_scene, use_mis):
    # The wavefront machinery (pool size, regeneration rounds, per-warp queue segments and their holes, paths kept
    # in registers vs parked in the queue) must not leak into the image: RNG counters are per (pixel, sample, depth),
    # so any pool gives the same frame up to the order of the fp32 atomic adds.
    g = gpu_scene("flying_unicorn")
    W, H, spp = 160, 120, 32
    ref = g.render(W, H, spp, seed=11, use_mis=use_mis).astype(int)
    st_ref = g.stats()
    for pool in (1024, 6000, 1 << 16):
        f = g.render(W, H, spp, seed=11, use_mis=use_mis, pool_paths=pool).astype(int)
        st = g.stats()
        assert np.abs(f - ref).max() <= 1, pool
        for k in ("samples", "rays_primary", "rays_extension", "rays_shadow", "rays_bvh", "shadow_bvh"):
            assert st[k] == st_ref[k], (pool, k)       # the same rays are traced, only their schedule differs
        assert st["iterations"] >= st_ref["iterations"]


def test_paths_stay_in_registers(gpu_scene):
    # cornell_box has no mesh: no ray ever needs k_traverse, so k_shade follows every path from the camera ray to its
    # end without queueing it again (one queue entry per sample, written by k_generate)
    g = gpu_scene("cornell_box")
    g.render(320, 240, 16, seed=4)
    st = g.stats()
    assert st["rays_bvh"] == 0 and st["shadow_bvh"] == 0
    assert st["paths_queued"] <= st["samples"] // 10     # only the paths parked at the tail of a large launch (< 20 per warp)
    g2 = gpu_scene("flying_unicorn")
    g2.render(320, 240, 16, seed=4)
    st2 = g2.stats()
    # a path is queued when its next ray can reach the mesh box (or its shadow ray could): far fewer than one entry per vertex
    assert st2["rays_bvh"] - st2["rays_primary"] <= st2["paths_queued"] < 0.35 * st2["rays_extension"]


@pytest.mark.parametrize("use_mis", [False, True])
def test_analytic_table_variants_agree(gpu_scene, monkeypatch, use_mis):
    # k_shade reads the plane / sphere table (a) from the kernel parameters with loops unrolled for the exact counts of the
    # reference scenes, (b) from the kernel parameters through a generic 8-slot form, (c) from shared memory in a loop.
    # Same arithmetic in the same order per ray -> the same frame (up to the order of the fp32 atomic adds).
    for name in ("cornell_box", "flying_unicorn"):
        g = gpu_scene(name)
        ref = g.render(200, 150, 16, seed=21, use_mis=use_mis).astype(int)
        st_ref = g.stats()
        for var in ("RTB_GENERIC_TABLE", "RTB_NO_SMALL_TABLE"):
            monkeypatch.setenv(var, "1")
            f = g.render(200, 150, 16, seed=21, use_mis=use_mis).astype(int)
            st = g.stats()
            monkeypatch.delenv(var)
            assert np.abs(f - ref).max() <= 1, (name, var)
            for k in ("rays_extension", "rays_shadow", "rays_bvh", "shadow_bvh"):
                assert st[k] == st_ref[k], (name, var, k)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_tile_shards_reassemble_the_frame(gpu_scene, rtb, world):
    # per-pixel RNG counters are independent of the sharding, so the union of the ranks' tiles is the
    # single-GPU frame (ranks emulated one after the other on one device; rtb_render leaves foreign tiles untouched)
    g = gpu_scene("flying_unicorn")
    W, H, spp = 200, 150, 16
    whole = g.render(W, H, spp, seed=5)
    out = np.zeros((H, W, 3), dtype=np.uint8)
    total = 0
    for r in range(world):
        g.render(W, H, spp, seed=5, rank=r, world=world, out=out)
        total += g.stats()["samples"]
    assert total == W * H * spp
    assert np.abs(out.astype(int) - whole.astype(int)).max() <= 1


def test_device_render_and_untile(gpu_scene, rtb):
    import torch

    from raytracer_server_b200 import _abi, sharding

    g = gpu_scene("cubes")
    W, H, spp, world = 600, 450, 8, 4
    stride = sharding.shard_stride(W, H, world)
    shards = torch.zeros(world * stride, dtype=torch.uint8, device="cuda:0")
    for r in range(world):
        p = rtb.make_params(W, H, spp, seed=6, rank=r, world=world)
        g.render_device(p, shards[r * stride:].data_ptr())
    frame = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda:0")
    p = rtb.make_params(W, H, spp, world=world)
    torch.cuda.synchronize()
    assert _abi.lib().rtb_untile_device(C.byref(p), C.c_void_p(shards.data_ptr()), stride, C.c_void_p(frame.data_ptr()), 0) == 0
    whole = g.render(W, H, spp, seed=6)
    assert np.abs(frame.cpu().numpy().astype(int) - whole.astype(int)).max() <= 1
    assert np.array_equal(sharding.untile_numpy(shards.cpu().numpy().reshape(world, stride), W, H, world), frame.cpu().numpy())


def test_subpixel_sums_explain_the_bytes(gpu_scene, rtb):
    # the resolve step is sample_pixel's tail + gamma_correct + `as u8` (src/server.rs:360-368, 187-189)
    import torch

    from raytracer_server_b200 import sharding

    g = gpu_scene("cornell_box")
    W, H, spp = 96, 64, 16
    n = sharding.local_pixels(W, H, 0, 1)
    rgb = torch.zeros(n * 3, dtype=torch.uint8, device="cuda:0")
    sub = torch.zeros(n * 4 * 4, dtype=torch.float32, device="cuda:0")
    g.render_device(rtb.make_params(W, H, spp, seed=1), rgb.data_ptr(), sub.data_ptr())
    s = sub.cpu().numpy().reshape(n, 4, 4)[:, :, :3].astype(np.float64)
    px = (np.clip(s, 0, 1) * 0.25).sum(axis=1)
    want = (np.clip(px, 0, 1) ** (1 / 2.2) * 255 + 0.5).astype(np.uint8)
    got = rgb.cpu().numpy().reshape(n, 3)
    xy = sharding.tile_map(W, H, 0, 1)
    ok = xy[:, 0] >= 0
    assert np.abs(got[ok].astype(int) - want[ok].astype(int)).max() <= 1
    assert (s[ok].max(axis=(1, 2)) > 1.0).any()      # some sub-pixels exceed 1: the per-sub-pixel clamp matters


def test_full_size_configs_properties(gpu_scene):
    # BASELINE.json configs[0..2] at full size: properties that need no oracle render
    for name, (W, H, spp) in (("cornell_box", (600, 450, 64)), ("cubes", (600, 450, 256)), ("flying_unicorn", (1920, 1080, 16))):
        g = gpu_scene(name)
        f = g.render(W, H, spp, seed=9)
        st = g.stats()
        assert st["samples"] == W * H * (spp // 4) * 4
        assert f.shape == (H, W, 3) and f.min() >= 0 and 60 < f.mean() < 180
        # left half red-ish wall, right half blue-ish wall, mirrored means of the grey channels agree roughly
        assert f[:, : W // 8, 0].mean() > f[:, : W // 8, 2].mean()
        assert f[:, -W // 8:, 2].mean() > f[:, -W // 8:, 0].mean()


def test_streaming_job_matches_blocking_render(gpu_scene, rtb):
    g = gpu_scene("cornell_box")
    W, H, spp = 130, 70, 8        # width not a multiple of 60: windows of 60, 60, 10 (src/server.rs:254-280)
    whole = g.render(W, H, spp, seed=4)
    job = rtb.RenderJob(g, W, H, spp, seed=4)
    frame = np.zeros((H, W, 3), dtype=np.uint8)
    order, sizes = [], set()
    for m in job.messages():
        assert m[0] == 0
        n, x, y = m[1], int.from_bytes(m[2:4], "little"), int.from_bytes(m[4:6], "little")
        assert len(m) == 6 + 3 * n
        frame[y, x: x + n] = np.frombuffer(m[6:], dtype=np.uint8).reshape(n, 3)
        order.append((y, x))
        sizes.add(n)
    assert job.close() is False
    assert order == sorted(order) and len(order) == H * 3 and sizes == {60, 10}
    assert np.abs(frame.astype(int) - whole.astype(int)).max() <= 1


def test_progressive_passes_converge(gpu_scene, rtb):
    g = gpu_scene("cornell_box")
    W, H, spp, passes = 120, 90, 64, 4
    final = g.render(W, H, spp, seed=8).astype(int)
    job = rtb.RenderJob(g, W, H, spp, seed=8, passes=passes)
    frames, frame, count = [], np.zeros((H, W, 3), dtype=np.uint8), 0
    per_pass = H * 2
    for m in job.messages():
        n, x, y = m[1], int.from_bytes(m[2:4], "little"), int.from_bytes(m[4:6], "little")
        frame[y, x: x + n] = np.frombuffer(m[6:], dtype=np.uint8).reshape(n, 3)
        count += 1
        if count % per_pass == 0:
            frames.append(frame.astype(int).copy())
    job.close()
    assert len(frames) == passes
    errs = [np.abs(f - final).mean() for f in frames]
    assert errs[-1] <= 0.01 and errs[0] > errs[-1]     # the last pass IS the final frame; earlier ones are noisier


def test_progressive_whole_frames(gpu_scene, rtb):
    # BASELINE config 5: 1 sample per pixel per frame, cycling the four sub-pixels (4 frames = reference spp 4)
    g = gpu_scene("cornell_box")
    W, H, spp = 200, 150, 32
    final = g.render(W, H, spp, seed=8).astype(int)
    job = rtb.RenderJob(g, W, H, spp, seed=8, passes=spp)
    got = list(job.frames())
    job.close()
    assert [i for i, _ in got] == list(range(spp))
    errs = [np.abs(f.astype(int) - final).mean() for _, f in got]
    assert errs[-1] <= 0.01 and errs[0] > errs[3] > errs[-1]
    # after the first frame only sub-pixel 0 has a sample: the frame is a quarter as bright as the final one at most
    assert got[0][1].mean() < 0.75 * final.mean()


def test_job_cancel(gpu_scene, rtb):
    g = gpu_scene("flying_unicorn")
    job = rtb.RenderJob(g, 1920, 1080, 4096, seed=1)   # seconds of work
    t = threading.Timer(0.2, job.stop)
    t.start()
    t0 = time.time()
    msgs = list(job.messages())
    assert job.close() is True and msgs == [] and time.time() - t0 < 5.0


def test_render_cancel_flag(gpu_scene, rtb):
    from raytracer_server_b200 import _abi

    g = gpu_scene("cubes")
    flag = C.c_int32(1)
    p = rtb.make_params(600, 450, 64)
    out = np.zeros((450, 600, 3), dtype=np.uint8)
    rc = _abi.lib().rtb_render(g._h, C.byref(p), out.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(flag))
    assert rc == _abi.RTB_ECANCELLED


def test_invalid_params(gpu_scene, rtb):
    g = gpu_scene("cubes")
    for kw in (dict(width=0, height=10, spp=4), dict(width=10, height=10, spp=-1), dict(width=70000, height=10, spp=4)):
        with pytest.raises(rtb.RtbError) as e:
            g.render(kw["width"], kw["height"], kw["spp"], out=np.zeros((1,), dtype=np.uint8) if False else None) if kw["width"] > 0 and kw["width"] < 65536 else g.trace_primary(kw["width"], kw["height"]) if kw["width"] == 0 else g.render(10, 10, 4, rank=3, world=2)
        assert e.value.code == -7


def test_concurrent_renders_share_a_scene(gpu_scene):
    # the reference shares Arc<HashMap<String, Scene>> between connections (src/server.rs:24): a scene handle is
    # immutable and every render call owns its stream + scratch
    g = gpu_scene("cubes")
    ref = g.render(200, 150, 16, seed=12).astype(int)
    results = [None] * 4

    def work(i):
        results[i] = g.render(200, 150, 16, seed=12).astype(int)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for r in results:
        assert np.abs(r - ref).max() <= 1


def test_fp32_peak_probe(rtb):
    assert 30.0 < rtb.fp32_peak_tflops(0) < 90.0
