"""Behaviour of the render entry points on the GPU: frame conventions, size-independent properties
at BASELINE.json sizes, tile sharding, streaming jobs, cancellation, error paths."""
import ctypes as C
import os
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


def test_paths_stay_in_registers(gpu_scene, monkeypatch):
    monkeypatch.setenv("RTB_INLINE_TAIL", "0")          # the queued form throughout (the inline tail would traverse in place)
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
    # a path is queued when its next ray can reach the mesh box (a queued shadow ray alone no longer parks it while the shadow
    # queue has room): far fewer than one entry per vertex
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


@pytest.mark.parametrize("workers", ["1", "2"])
def test_streaming_bands_match_blocking_render(gpu_scene, rtb, monkeypatch, workers):
    # a streaming job renders the frame in bands of tile rows and publishes each band from a side stream while the
    # next ones render (RenderJob::run sends windows as they are sampled, src/server.rs:166-194); forced to 1 tile
    # row per band here: 6 bands, the last one partial (170 = 5 * 32 + 10 rows), one or two worker threads
    monkeypatch.setenv("RTB_BAND_TILE_ROWS", "1")
    monkeypatch.setenv("RTB_JOB_WORKERS", workers)
    g = gpu_scene("flying_unicorn")
    W, H, spp = 200, 170, 16
    whole = g.render(W, H, spp, seed=21).astype(int)
    job = rtb.RenderJob(g, W, H, spp, seed=21)
    frame, order = np.zeros((H, W, 3), dtype=np.uint8), []
    for m in job.messages():
        n, x, y = m[1], int.from_bytes(m[2:4], "little"), int.from_bytes(m[4:6], "little")
        frame[y, x: x + n] = np.frombuffer(m[6:], dtype=np.uint8).reshape(n, 3)
        order.append((y, x))
    st = job.stats()
    assert job.close() is False
    assert order == sorted(order) and len(order) == H * 4          # windows of 60, 60, 60, 20
    assert np.abs(frame.astype(int) - whole).max() <= 1
    assert st["samples"] == W * H * spp and 0 < st["first_record_ms"] <= st["wall_ms"]
    assert job.stats()["samples"] == W * H * spp                    # the final counters survive close()
    # the whole-frame form waits for the last band
    job = rtb.RenderJob(g, W, H, spp, seed=21)
    (idx, f), = list(job.frames())
    job.close()
    assert idx == 0 and np.abs(f.astype(int) - whole).max() <= 1


def test_first_record_arrives_while_the_frame_renders(gpu_scene, rtb, parity_log):
    # BASELINE configs[2]: 1920x1080, 256 spp.  Bands of >= 8 Mi samples: the first rows are out long before the frame is done
    g = gpu_scene("flying_unicorn")
    W, H, spp = 1920, 1080, 256
    g.render(W, H, 8)
    t0 = time.time()
    g.render(W, H, spp, seed=5)
    blocking_ms = (time.time() - t0) * 1e3
    for attempt in range(2):    # the first job on a scene allocates its band-sized contexts and pinned frame (100s of ms, once)
        job = rtb.RenderJob(g, W, H, spp, seed=5)
        n = sum(1 for _ in job.messages())
        st = job.stats()
        job.close()
    assert n == H * 32 and st["samples"] == W * H * spp
    parity_log("gpu/streaming/flying_unicorn_1920x1080_256spp", first_record_ms=st["first_record_ms"], job_wall_ms=st["wall_ms"],
               blocking_render_ms=blocking_ms, msamples_per_s=st["samples"] / st["wall_ms"] / 1e3)
    assert st["first_record_ms"] <= 0.10 * st["wall_ms"]
    assert st["wall_ms"] <= 1.35 * blocking_ms


def test_sample_pixels_is_the_frame(gpu_scene, rtb):
    # sample_pixel (src/server.rs:320-364) through rtb_sample_pixels: same random numbers as the frame render, so the
    # truncated Vec3 is the frame's byte (up to the order of the fp32 adds); only the listed pixels are traced
    from raytracer_server_b200.host import sample_pixel, sample_pixels

    g = gpu_scene("flying_unicorn")
    W, H, spp = 160, 120, 16
    frame = g.render(W, H, spp, seed=6).astype(int)
    rng = np.random.default_rng(1)
    xs, ys = rng.integers(0, W, 500), rng.integers(0, H, 500)
    v = sample_pixels(xs, ys, W, H, spp, g, seed=6)
    st = g.stats()
    assert st["samples"] == 500 * spp
    assert v.shape == (500, 3) and v.min() >= 0.5 and v.max() <= 255.5
    assert np.abs(np.floor(v).astype(int) - frame[ys, xs]).max() <= 1
    one = sample_pixel(17, H - 40 - 1, W, H, spp, g, seed=6)       # the reference's signature: bottom-up y
    assert np.abs(np.floor(one).astype(int) - frame[40, 17]).max() <= 1
    assert (sample_pixels([3], [3], W, H, 3, g) == 0.5).all()       # spp < 4: no samples, gamma_correct(0) = 0.5


def test_stats_belong_to_the_calling_thread(gpu_scene):
    g = gpu_scene("cubes")
    out = {}

    def work(name, w, h, spp):
        g.render(w, h, spp, seed=1)
        out[name] = g.stats()["samples"]

    ts = [threading.Thread(target=work, args=("a", 64, 48, 8)), threading.Thread(target=work, args=("b", 96, 64, 16))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert out == {"a": 64 * 48 * 8, "b": 96 * 64 * 16}


def test_bvh_depth_is_checked_against_the_stack(rtb, gpu_scene, monkeypatch):
    assert 8 <= gpu_scene("flying_unicorn").info.bvh_depth <= 96 and gpu_scene("cornell_box").info.bvh_depth == 0
    monkeypatch.setenv("RTB_BVH_MAX_DEPTH", "6")     # pretend the traversal stack holds 6 levels: PLOC and Karras trees are both deeper
    with pytest.raises(rtb.LoadTomlError) as e:
        rtb.Scene.from_toml(scene_path("flying_unicorn"), device=0)
    assert e.value.kind == "Unsupported" and "levels deep" in str(e.value)
    monkeypatch.setenv("RTB_BVH_MAX_DEPTH", "96")
    assert rtb.Scene.from_toml(scene_path("cubes"), device=0).info.bvh_depth >= 1


def test_balance_heuristic_mis_mode(gpu_scene, rtb, parity_log):
    # RTB_EST_MIS_BALANCE: the estimator the reference's TODO (src/scene.rs:187) asks for — NOT a parity mode.
    # Same expectation as the live NEE estimator, less variance where BRDF sampling finds the light easily (the
    # ceiling right above the lamp), and the dead "MIS" branch stays what it was.
    g = gpu_scene("cornell_box")
    W, H, spp, n = 600, 450, 4096, 400_000
    rng = np.random.default_rng(3)
    px, py, si = rng.integers(0, W, n), rng.integers(0, H, n), rng.integers(0, spp, n)
    nee = g.sample_radiance(W, H, spp, px, py, si, seed=1).astype(np.float64)
    bal = g.sample_radiance(W, H, spp, px, py, si, seed=1, estimator=rtb.EST_MIS_BALANCE).astype(np.float64)
    se = np.sqrt(nee.var(0) / n + bal.var(0) / n)
    parity_log("gpu/mis_balance/cornell_box", paths=n, mean_nee=nee.mean(0), mean_balance=bal.mean(0), std_nee=nee.std(0), std_balance=bal.std(0))
    assert (np.abs(nee.mean(0) - bal.mean(0)) < 4 * se).all(), (nee.mean(0), bal.mean(0), se)
    assert not np.array_equal(nee, bal)
    # the hot spot on the ceiling above the light (screen rows 30..60, columns 250..350): NEE suffers from 1 / r^2 there
    m = 100_000
    px, py, si = rng.integers(250, 350, m), rng.integers(30, 60, m), rng.integers(0, spp, m)
    v_nee = g.sample_radiance(W, H, spp, px, py, si, seed=2).astype(np.float64)
    v_bal = g.sample_radiance(W, H, spp, px, py, si, seed=2, estimator=rtb.EST_MIS_BALANCE).astype(np.float64)
    parity_log("gpu/mis_balance/cornell_box_ceiling", paths=m, mean_nee=v_nee.mean(0), mean_balance=v_bal.mean(0), std_nee=v_nee.std(0), std_balance=v_bal.std(0))
    assert (v_bal.std(0) < 0.93 * v_nee.std(0)).all()          # measured: 0.86 (the path's deeper bounces carry most of the variance)
    se = np.sqrt(v_nee.var(0) / m + v_bal.var(0) / m)
    assert (np.abs(v_nee.mean(0) - v_bal.mean(0)) < 4 * se).all()
    # frames: the whole-frame entry point accepts the mode; the scene's mesh-light twin refuses it
    f = g.render(200, 150, 64, seed=5, estimator=rtb.EST_MIS_BALANCE).astype(int)
    assert abs(f.mean() - g.render(200, 150, 64, seed=6).astype(int).mean()) < 0.02 * f.mean()


def test_binning_knob_changes_nothing_but_the_order(gpu_scene):
    # coherence binning of the LBVH rays (tuning[3], off by default; DESIGN.md "Ray sorting, measured"): k_traverse fetches its
    # rays through a permutation, results and ray counts are those of the unbinned run
    g = gpu_scene("flying_unicorn")
    W, H, spp = 240, 180, 32
    ref = g.render(W, H, spp, seed=4).astype(int)
    base = g.stats()
    for bits, octm in ((3, False), (5, True)):
        f = g.render(W, H, spp, seed=4, bin_bits=bits, bin_octant_major=octm).astype(int)
        st = g.stats()
        assert np.abs(f - ref).max() <= 1
        assert all(st[k] == base[k] for k in ("samples", "rays_primary", "rays_extension", "rays_shadow", "rays_bvh", "shadow_bvh"))
        assert st["kernel_launches"] >= 7 * st["iterations"]      # k_bin_keys, _scan, _scatter on top of the four per iteration


def test_streaming_job_in_reference_octree_mode(gpu_scene, rtb, monkeypatch):
    # accel travels with the job: bands rendered through the reference's octrees equal the blocking render in that mode,
    # and differ from the LBVH image where the octree's early exit picks other triangles (flying_unicorn)
    monkeypatch.setenv("RTB_BAND_TILE_ROWS", "2")
    g = gpu_scene("flying_unicorn")
    W, H, spp = 200, 150, 16
    octree = g.render(W, H, spp, seed=9, accel=rtb.ACCEL_OCTREE_REFERENCE).astype(int)
    lbvh = g.render(W, H, spp, seed=9).astype(int)
    job = rtb.RenderJob(g, W, H, spp, seed=9, accel=rtb.ACCEL_OCTREE_REFERENCE)
    (idx, f), = list(job.frames())
    job.close()
    assert np.abs(f.astype(int) - octree).max() <= 1
    assert (np.abs(octree - lbvh) > 8).mean() > 0.002
    assert g.info.octree_nodes == 47183 and g.info.octree_tri_refs == 187766     # built on first use (SURVEY's census of the reference structure)
    assert rtb.Scene.from_toml(scene_path("cubes"), device=0).render(64, 48, 8, accel=rtb.ACCEL_OCTREE_REFERENCE).shape == (48, 64, 3)


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


def test_every_mode_with_tiny_pools(monkeypatch):
    # tools/gpu_sanitize.py: every kernel, estimator, accel mode, probe / pixel-list entry point, the binning knob, banded,
    # progressive and cancelled jobs on pools of 2-4 k paths (many regeneration rounds, queue segments with holes).
    # Written for compute-sanitizer (closed on this pool); here it must simply run clean — the library reports queue
    # overflows, CUDA errors and inconsistent counters as errors.
    import runpy

    from conftest import ROOT

    for k in ("RTB_NO_GRAPH", "RTB_BAND_TILE_ROWS"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv("RTB_KEEP", "1")       # (monkeypatch restores os.environ after the script's own changes)
    runpy.run_path(os.path.join(ROOT, "tools", "gpu_sanitize.py"), run_name="__main__")


def test_imported_scene_is_bit_identical(rtb, gpu_scene):
    # the scene + BVH broadcast of a multi-GPU job (sharding.broadcast_scene): a handle imported from another handle's export
    # holds the same LBVH tables byte for byte, so traversal results are IDENTICAL (not merely close) and no build runs
    a = gpu_scene("flying_unicorn")
    blob = a.export()
    b = rtb.Scene.from_export(blob, device=0)
    assert (b.info.bvh_nodes, b.info.bvh_leaves, b.info.bvh_depth, b.info.n_triangles) == (a.info.bvh_nodes, a.info.bvh_leaves, a.info.bvh_depth, 37380)
    assert np.array_equal(b.export(), blob)                          # export(import(x)) == x
    rng = np.random.default_rng(2)
    n = 100_000
    lo, hi = np.array(a.info.bvh_min), np.array(a.info.bvh_max)
    org = (lo + (hi - lo) * (rng.random((n, 3)) * 1.4 - 0.2)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    ra, rb = a.trace_rays(org, d), b.trace_rays(org, d)
    assert all(np.array_equal(ra[k], rb[k]) for k in ("obj", "tri", "t")) and (ra["tri"] >= 0).mean() > 0.05
    assert np.abs(a.render(160, 120, 16, seed=3).astype(int) - b.render(160, 120, 16, seed=3).astype(int)).max() <= 1
    oa, ob = (s.render(96, 72, 8, seed=3, accel=rtb.ACCEL_OCTREE_REFERENCE).astype(int) for s in (a, b))
    assert np.abs(oa - ob).max() <= 1                                  # the octrees are rebuilt from the shipped f64 vertices
    c = rtb.Scene.from_export(rtb.Scene.from_toml(scene_path("cubes"), device=-1).export(), device=0)   # objects only: the importer builds
    assert c.info.bvh_nodes > 0 and np.abs(c.render(64, 48, 8, seed=1).astype(int) - gpu_scene("cubes").render(64, 48, 8, seed=1).astype(int)).max() <= 1


@pytest.mark.parametrize("name,use_mis", [("flying_unicorn", False), ("flying_unicorn", True), ("cubes", False)])
def test_inline_tail_changes_nothing_but_the_iterations(gpu_scene, monkeypatch, name, use_mis):
    # Once few paths are left (RTB_INLINE_TAIL, default 2 Mi) k_shade traverses the LBVH itself and the run ends in one
    # launch instead of ~100 nearly empty iterations.  Same rays, same frame; far fewer iterations.
    g = gpu_scene(name)
    W, H, spp = 320, 240, 64
    monkeypatch.setenv("RTB_INLINE_TAIL", "0")
    ref = g.render(W, H, spp, seed=17, use_mis=use_mis).astype(int)
    base = g.stats()
    for tail in ("1000", "2097152"):
        monkeypatch.setenv("RTB_INLINE_TAIL", tail)
        f = g.render(W, H, spp, seed=17, use_mis=use_mis).astype(int)
        st = g.stats()
        assert np.abs(f - ref).max() <= 1
        assert all(st[k] == base[k] for k in ("samples", "rays_primary", "rays_extension", "rays_shadow", "rays_bvh", "shadow_bvh"))
        assert st["iterations"] < base["iterations"]
    assert st["iterations"] <= 0.5 * base["iterations"]
