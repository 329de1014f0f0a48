"""Pins for the CPU oracle (oracle/rt_oracle.cpp).

The reference holds exactly one unit test for this path — `test_octants`
(src/geometry.rs:1115-1131) — reproduced first.  Everything else is a hand-derived known answer
for a rule quoted from the reference source (file:line in each test), because the Rust reference
cannot be built here and its renders are unseeded.
"""
import math
import os

import numpy as np
import pytest

from conftest import NCPU, scene_path


def scene_from(O, objects, cam_pos=(0, 0, 10), cam_dir=(0, 0, -1)):
    txt = f"[camera]\npos = {list(map(float, cam_pos))}\ndir = {list(map(float, cam_dir))}\n"
    for ob in objects:
        txt += "\n[[objects]]\n" + ob + "\n"
    return O.OracleScene.from_toml_string(txt, os.path.dirname(scene_path("x")) + "/assets")


LIGHT = 'emitted = [10.0, 10.0, 10.0]\nbrdf = { type = "diffuse", kd = [0.0, 0.0, 0.0] }\ngeometry = { type = "sphere", pos = [0.0, 100.0, 0.0], r = 1.0 }'


def test_octants_reference_unit_test(oracle_mod):
    # src/geometry.rs:1115-1131, verbatim expectations (bit 2 = x, bit 1 = y, bit 0 = z)
    got = oracle_mod.octants([-1, -1, -1], [1, 1, 1])
    want = [((-1, -1, -1), (0, 0, 0)), ((-1, -1, 0), (0, 0, 1)), ((-1, 0, -1), (0, 1, 0)), ((-1, 0, 0), (0, 1, 1)),
            ((0, -1, -1), (1, 0, 0)), ((0, -1, 0), (1, 0, 1)), ((0, 0, -1), (1, 1, 0)), ((0, 0, 0), (1, 1, 1))]
    assert np.array_equal(got, np.array(want, dtype=float))


def test_philox_known_answers(oracle_mod):
    # Random123 kat_vectors, philox4x32-10
    P = oracle_mod.philox4x32_10
    assert P([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert P([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert P([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_sphere_hits_and_eps(oracle_mod):
    # src/geometry.rs:514-550: near root if > 1e-4, else far root if > 1e-4; normal faces the ray; pos not offset
    sc = scene_from(oracle_mod, ['brdf = { type = "diffuse", kd = [0.5,0.5,0.5] }\ngeometry = { type = "sphere", pos = [0.0,0.0,0.0], r = 2.0 }', LIGHT])
    r = sc.trace_rays([[0, 0, 10], [0, 0, 0], [0, 0, 2.0], [0, 0, 2.0 - 5e-5], [0, 5, 10]],
                      [[0, 0, -1], [0, 0, 1], [0, 0, 1], [0, 0, -1], [0, 0, -1]], want_geom=True)
    assert r["obj"].tolist() == [0, 0, -1, 0, -1]
    assert r["t"][0] == pytest.approx(8.0, abs=1e-12) and r["t"][1] == pytest.approx(2.0, abs=1e-12)
    assert r["t"][3] == pytest.approx(4.0 - 5e-5, abs=1e-9)       # near root 5e-5 < eps -> far side
    assert np.allclose(r["n"][0], [0, 0, 1]) and np.allclose(r["n"][1], [0, 0, -1])  # flipped toward the ray
    assert np.allclose(r["pos"][0], [0, 0, 2.0])


def test_plane_rules(oracle_mod):
    # src/geometry.rs:551-568: |d.n| < 1e-4 -> miss; t >= 0 (no eps); pos offset +1e-5 * facing normal
    sc = scene_from(oracle_mod, ['brdf = { type = "diffuse", kd = [0.5,0.5,0.5] }\ngeometry = { type = "plane", pos = [0.0,0.0,0.0], n = [0.0,1.0,0.0] }', LIGHT])
    s = 9.9e-5
    d_par = [math.sqrt(1 - s * s), -s, 0.0]
    s2 = 1.01e-4
    d_ok = [math.sqrt(1 - s2 * s2), -s2, 0.0]
    r = sc.trace_rays([[0, 1, 0], [0, 1, 0], [0, 0, 0], [0, -1, 0], [0, 1, 0]],
                      [[0, -1, 0], d_par, [0, -1, 0], [0, 1, 0], d_ok], want_geom=True)
    assert r["obj"].tolist() == [0, -1, 0, 0, 0]
    assert r["t"][0] == 1.0 and r["t"][2] == 0.0            # t == 0 is a hit
    assert np.allclose(r["pos"][0], [0, 1e-5, 0]) and np.allclose(r["n"][0], [0, 1, 0])
    assert np.allclose(r["pos"][3], [0, -1e-5, 0]) and np.allclose(r["n"][3], [0, -1, 0])  # from below: flipped


def test_triangle_rules(oracle_mod):
    # src/geometry.rs:637-670 on the top face of a unit cube (prism triangles 2,3,7 / 2,6,7; y = 1)
    sc = scene_from(oracle_mod, ['brdf = { type = "diffuse", kd = [0.5,0.5,0.5] }\ngeometry = { type = "cube", pos = [0.0,0.0,0.0], size = 1.0 }', LIGHT])
    r = sc.trace_rays([[0.25, 3, 0.5], [0.25, 1 + 5e-5, 0.5], [0.25, 1 + 2e-4, 0.5], [2, 3, 0.5]],
                      [[0, -1, 0], [0, -1, 0], [0, -1, 0], [0, -1, 0]], want_geom=True)
    assert r["obj"].tolist() == [0, 0, 0, -1]
    assert r["t"][0] == pytest.approx(2.0, abs=1e-12)
    assert np.allclose(r["n"][0], [0, 1, 0]) and np.allclose(r["pos"][0], [0.25, 1 + 1e-5, 0.5])
    assert r["t"][1] == pytest.approx(1 + 5e-5, abs=1e-9)   # top face at t = 5e-5 <= 1e-4 rejected -> bottom face
    assert r["t"][2] == pytest.approx(2e-4, abs=1e-9)       # t = 2e-4 > 1e-4 accepted
    # triangle index is the position in Mesh::prism's index list (src/geometry.rs:853-860): top = 8, 9
    assert r["tri"][0] in (8, 9)


def test_trace_ray_lowest_index_wins_ties(oracle_mod, oracle_scene):
    # src/scene.rs:277-284 strict '<': object 5 duplicates object 1's plane in every reference scene
    # (scenes/cornell_box.toml objects 1 and 5) and must never be reported
    for name in ("cornell_box", "cubes", "flying_unicorn"):
        sc = oracle_scene(name)
        org, dirs = sc.primary_rays(80, 60)
        rng = np.random.default_rng(0)
        d2 = rng.normal(size=(4000, 3))
        d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
        o2 = np.tile([[50.0, 40.0, 150.0]], (4000, 1))
        r = sc.trace_rays(np.vstack([org, o2]), np.vstack([dirs, d2]))
        assert (r["obj"] != 5).all()
        assert (r["obj"] == 1).any()
        assert (r["obj"] >= 0).mean() > 0.99   # infinite corridor: essentially nothing misses


def test_light_source_rule_and_errors(oracle_mod):
    # src/scene.rs:129-137: first object with any |emitted| >= 1e-5
    dim = 'emitted = [0.000009, 0.0, 0.0]\nbrdf = { type = "diffuse", kd = [0.5,0.5,0.5] }\ngeometry = { type = "sphere", pos = [5.0,0.0,0.0], r = 1.0 }'
    sc = scene_from(oracle_mod, [dim, LIGHT])
    assert sc.light_source == 1
    with pytest.raises(oracle_mod.OracleError):
        scene_from(oracle_mod, [dim])


def test_unicorn_octree_census(oracle_mod, oracle_scene):
    # SURVEY §2.2 probe of Octree::build on the transformed mesh (src/geometry.rs:1149-1216)
    sc = oracle_scene("flying_unicorn")
    st = sc.mesh_stats(6)
    assert st["triangles"] == 37380
    assert (st["octree_parents"], st["octree_leaves"], st["octree_tri_refs"]) == (9540, 37643, 187766)
    assert np.allclose(st["bbox_min"], [11.657, -2.801, 56.176], atol=1e-3)
    assert np.allclose(st["bbox_max"], [52.060, 59.347, 91.408], atol=1e-3)


def test_exact_bvh_equals_brute_force(oracle_mod):
    # the 'exact' accelerator must return what Mesh::intersect's linear scan returns (src/geometry.rs:887-903)
    sc = scene_from(oracle_mod, ['brdf = { type = "diffuse", kd = [0.5,0.5,0.5] }\ngeometry = { type = "mesh", path = "chair.obj" }', LIGHT])
    tris = sc.mesh_triangles(0)
    rng = np.random.default_rng(3)
    lo, hi = tris.reshape(-1, 3).min(0), tris.reshape(-1, 3).max(0)
    org = rng.uniform(lo - 1, hi + 1, size=(600, 3))
    tgt = rng.uniform(lo, hi, size=(600, 3))
    d = tgt - org
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = sc.trace_rays(org, d)
    # numpy brute force with the reference's own formulas
    a, b, c = tris[:, 0], tris[:, 1], tris[:, 2]
    n = np.cross(c - a, b - a)
    n /= np.linalg.norm(n, axis=1, keepdims=True)

    def det3(v0, v1, v2):
        return (v0[..., 0] * (v1[..., 1] * v2[..., 2] - v1[..., 2] * v2[..., 1]) - v1[..., 0] * (v0[..., 1] * v2[..., 2] - v0[..., 2] * v2[..., 1])
                + v2[..., 0] * (v0[..., 1] * v1[..., 2] - v0[..., 2] * v1[..., 1]))

    for i in range(org.shape[0]):
        md = -d[i][None, :]
        ab, ac, bb = b - a, c - a, org[i][None, :] - a
        with np.errstate(divide="ignore", invalid="ignore"):
            det = det3(md, ab, ac)
            t = det3(bb, ab, ac) / det
            u = det3(md, bb, ac) / det
            v = det3(md, ab, bb) / det
        ok = (np.abs(n @ d[i]) >= 1e-4) & ~((u < 0) | (u > 1) | (v < 0) | (u + v > 1)) & (t > 1e-4)
        if ok.any():
            k = np.flatnonzero(ok)[np.argmin(t[ok])]
            assert r["obj"][i] == 0 and r["tri"][i] == k and r["t"][i] == pytest.approx(t[k], rel=1e-12)
        else:
            assert r["obj"][i] != 0


def test_octree_faithful_vs_exact_census(oracle_mod, oracle_scene):
    # F6: the reference's octree is an early-exit structure, not nearest-hit.  Primary rays almost always
    # agree with the true nearest hit; the census is recorded, not gated tightly.
    sc = oracle_scene("flying_unicorn")
    org, dirs = sc.primary_rays(160, 120)
    exact = sc.trace_rays(org, dirs)
    sc.set_modes(oracle_mod.ACCEL_OCTREE_FAITHFUL, oracle_mod.EST_NEE)
    oct_ = sc.trace_rays(org, dirs, count_work=True)
    sc.set_modes(oracle_mod.ACCEL_EXACT, oracle_mod.EST_NEE)
    mesh = exact["obj"] == 6
    assert mesh.sum() > 1000
    agree = (oct_["obj"][mesh] == 6) & (oct_["tri"][mesh] == exact["tri"][mesh])
    assert agree.mean() > 0.99
    assert (oct_["obj"] == exact["obj"]).mean() > 0.999


def test_direct_light_closed_form(oracle_mod):
    # live estimator src/scene.rs:217-229 on a floor under a sphere light: a Lambertian point straight below a
    # uniform sphere emitter (radius r, centre distance d) reflects L = kd * Le * (r/d)^2; upward bounces leave
    # the scene or hit the light (whose emission the diffuse branch never adds), so that is the whole radiance.
    kd, Le, r, d = 0.5, 10.0, 1.0, 10.0
    floor = f'brdf = {{ type = "diffuse", kd = [{kd},{kd},{kd}] }}\ngeometry = {{ type = "plane", pos = [0.0,0.0,0.0], n = [0.0,1.0,0.0] }}'
    light = f'emitted = [{Le},{Le},{Le}]\nbrdf = {{ type = "diffuse", kd = [0.0,0.0,0.0] }}\ngeometry = {{ type = "sphere", pos = [0.0,{d},0.0], r = {r} }}'
    sc = scene_from(oracle_mod, [floor, light], cam_pos=(0, 5, 20), cam_dir=(0, -0.25, -1))
    W = H = 101
    n = 4000
    L = sc.sample_radiance(W, H, 4 * n, 5, np.full(n, 50), np.full(n, 50), np.arange(n))
    # the centre pixel looks at (0,0,0) up to a footprint of ~0.1 units
    assert L.mean(axis=0) == pytest.approx([kd * Le * (r / d) ** 2] * 3, rel=0.02)


def test_render_counts_and_spp_rule(oracle_mod, oracle_scene):
    # src/server.rs:332: num_samples = spp / 4 -> spp < 4 renders black; effective spp = 4 * (spp / 4)
    sc = oracle_scene("cornell_box")
    r = sc.render(16, 12, 3)
    assert r["samples"] == 0 and (r["rgb8"] == 0).all()
    r = sc.render(16, 12, 7, nthreads=2)
    assert r["samples"] == 16 * 12 * 4
    a = sc.render(24, 18, 8, seed=1, nthreads=1)
    b = sc.render(24, 18, 8, seed=1, nthreads=-NCPU)
    assert np.array_equal(a["rgb8"], b["rgb8"])           # counter-based RNG: thread layout is irrelevant
    c = sc.render(24, 18, 8, seed=2)
    assert not np.array_equal(a["rgb8"], c["rgb8"])
    assert 25 < a["rays"] / a["samples"] < 35             # about 30 trace_ray calls per sample


def test_example_png_channel_means(oracle_mod, oracle_scene):
    # weak fixture: examples/cornell_box.png, examples/cubes.png (older revision, 64 spp) have channel means
    # (119.75, 97.58, 119.92) and (117.53, 96.22, 117.77).  Sanity only: several per cent.
    for name, want in (("cornell_box", (119.75, 97.58, 119.92)), ("cubes", (117.53, 96.22, 117.77))):
        r = oracle_scene(name).render(120, 90, 32, seed=4, nthreads=-NCPU)
        got = r["rgb8"].reshape(-1, 3).mean(axis=0)
        assert np.allclose(got, want, rtol=0.08), (name, got)
