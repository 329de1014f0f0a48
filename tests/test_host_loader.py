"""Host logic of the boundary, no GPU: the C++ TOML-subset + OBJ loader (device = -1 handles) against
Python's tomllib and the oracle's own set-up code, and the reference's error behaviour."""
import os
import tomllib

import numpy as np
import pytest

from conftest import SCENE_NAMES, SCENES, scene_path

ASSETS = os.path.join(SCENES, "assets")


def load(rtb, text, assets=ASSETS):
    return rtb.Scene.from_toml_string(text, assets_dir=assets, device=-1)


@pytest.mark.parametrize("name", SCENE_NAMES)
def test_loader_matches_tomllib_and_oracle(rtb, oracle_scene, name):
    with open(scene_path(name), "rb") as f:
        spec = tomllib.load(f)
    sc = rtb.Scene.from_toml(scene_path(name), device=-1)
    osc = oracle_scene(name)
    assert sc.info.n_objects == len(spec["objects"]) == osc.num_objects
    assert sc.light_source == osc.light_source == 8
    assert list(sc.info.camera_pos) == pytest.approx(spec["camera"]["pos"])
    assert list(sc.info.camera_dir) == pytest.approx(spec["camera"]["dir"], rel=1e-6)
    tri_total = 0
    for i, ob in enumerate(spec["objects"]):
        info = sc.object(i)
        assert info["emitted"] == ob.get("emitted", [0.0, 0.0, 0.0])
        assert info["brdf"] == {"diffuse": 0, "specular": 1, "phong": 2}[ob["brdf"]["type"]]
        gt = ob["geometry"]["type"]
        assert info["geometry"] == {"sphere": 0, "plane": 1, "cube": 2, "prism": 2, "mesh": 2}[gt]
        if gt == "sphere":
            assert info["pos"] == ob["geometry"]["pos"] and info["r"] == ob["geometry"]["r"]
        if gt == "plane":
            assert info["pos"] == ob["geometry"]["pos"] and info["n"] == ob["geometry"]["n"]
        if info["geometry"] == 2:
            st = osc.mesh_stats(i)
            assert info["n_triangles"] == st["triangles"] and info["first_triangle"] == tri_total
            assert np.allclose(info["bb_min"], st["bbox_min"], rtol=0, atol=1e-12)
            assert np.allclose(info["bb_max"], st["bbox_max"], rtol=0, atol=1e-12)
            want = osc.mesh_triangles(i).astype(np.float32)        # f64 set-up, rounded once to fp32
            got = sc.triangles()[tri_total: tri_total + info["n_triangles"]]
            assert np.array_equal(got, want)
            tri_total += info["n_triangles"]
    assert sc.info.n_triangles == tri_total


def test_transform_semantics(rtb, oracle_mod):
    # transforms apply in TOML order, about the bbox centre, with the scale() bbox quirk
    # (src/geometry.rs:445-510): translate then scale != scale then translate for the pivot
    def scene(tr):
        return f"""
[camera]
pos = [0.0, 0.0, 10.0]
dir = [0.0, 0.0, -1.0]
[[objects]]
brdf = {{ type = "diffuse", kd = [0.5, 0.5, 0.5] }}
geometry = {{ type = "prism", pos = [1.0, 2.0, 3.0], size = [1.0, 2.0, 4.0] }}
transforms = [ {tr} ]
[[objects]]
emitted = [1.0, 1.0, 1.0]
brdf = {{ type = "diffuse", kd = [0.0, 0.0, 0.0] }}
geometry = {{ type = "sphere", pos = [0.0, 9.0, 0.0], r = 1.0 }}
transforms = [ {{ scale = 2.0 }}, {{ translate = [1.0, 0.0, 0.0] }}, {{ rotate_x = 1.0 }} ]
"""
    for tr in ("{ scale = 2.0 }, { translate = [1.0, 0.5, 0.0] }, { rotate_y = 0.3 }",
               "{ rotate_z = -0.4 }, { scale = 0.5 }, { scale = 3.0 }, { rotate_x = 1.2 }",
               "{ translate = [5.0, 0.0, 0.0] }, { scale = 2.0 }, { scale = 2.0 }"):
        sc = load(rtb, scene(tr))
        osc = oracle_mod.OracleScene.from_toml_string(scene(tr), ASSETS)
        st = osc.mesh_stats(0)
        info = sc.object(0)
        assert np.allclose(info["bb_min"], st["bbox_min"], atol=1e-12) and np.allclose(info["bb_max"], st["bbox_max"], atol=1e-12)
        assert np.array_equal(sc.triangles(), osc.mesh_triangles(0).astype(np.float32))
        light = sc.object(1)
        assert light["r"] == 2.0 and light["pos"] == [1.0, 9.0, 0.0]      # spheres: scale r, translate pos, rotations ignored
    # the quirk itself: after one scale the stored box is min + (min - c) * s, wider than the vertices
    sc = load(rtb, scene("{ scale = 2.0 }"))
    info = sc.object(0)
    assert info["bb_min"] == [1.0 + (1.0 - 1.5) * 2, 2.0 + (2.0 - 3.0) * 2, 3.0 + (3.0 - 5.0) * 2]
    assert sc.triangles().reshape(-1, 3).min(0).tolist() == [0.5, 1.0, 1.0]


def test_plane_rotation_and_cube(rtb):
    sc = load(rtb, """
[camera]
pos = [0, 0, 10]     # integers are accepted where the reference's serde structs want f64
dir = [0, 0, -1]
[[objects]]
brdf = { type = "diffuse", kd = [0.5, 0.5, 0.5] }
geometry = { type = "plane", pos = [0.0, 0.0, 0.0], n = [0.0, 1.0, 0.0] }
transforms = [ { rotate_x = 1.5707963267948966 }, { translate = [0.0, 1.0, 0.0] }, { scale = 7.0 } ]
[[objects]]
emitted = [3, 3, 3]
brdf = { type = "specular", ks = [0.9, 0.9, 0.9] }
geometry = { type = "cube", pos = [0.0, 0.0, 0.0], size = 2.0 }
""")
    p = sc.object(0)
    assert p["n"] == pytest.approx([0.0, 0.0, 1.0], abs=1e-15) and p["pos"] == [0.0, 1.0, 0.0]
    c = sc.object(1)
    assert c["n_triangles"] == 12 and c["surface_area"] == pytest.approx(24.0) and c["brdf"] == 1
    assert sc.light_source == 1


def test_phong_and_toml_forms(rtb):
    sc = load(rtb, """
# dotted keys, literal strings, multi-line arrays, trailing commas, comments everywhere
camera.pos = [ 1.0,
               2.0,   # y
               3.0, ]
camera.dir = [0.0, 0.0, -1.0]

[[objects]]
emitted = [1e1, 1.0e+1, 1_0.0]
geometry = { type = 'sphere', pos = [0.0, 5.0, 0.0], r = 1.5 }
[objects.brdf]
type = "phong"
kd = 0.3
ks = 0.6
power = 20
color_d = [1.0, 0.5, 0.25]
color_s = [1.0, 1.0, 1.0]
""")
    o = sc.object(0)
    assert o["brdf"] == 2 and o["k"] == [0.3, 0.6, 20.0] and o["color_d"] == [1.0, 0.5, 0.25]
    assert o["emitted"] == [10.0, 10.0, 10.0] and list(sc.info.camera_pos) == [1.0, 2.0, 3.0]


LIGHT = """
[[objects]]
emitted = [1.0, 1.0, 1.0]
brdf = { type = "diffuse", kd = [0.0, 0.0, 0.0] }
geometry = { type = "sphere", pos = [0.0, 9.0, 0.0], r = 1.0 }
"""
CAM = "[camera]\npos = [0.0, 0.0, 10.0]\ndir = [0.0, 0.0, -1.0]\n"


@pytest.mark.parametrize("text,kind", [
    ("[camera]\npos = [0.0, 0.0, 0.0]\n", "Parse"),                                    # missing field `dir`
    (CAM, "Parse"),                                                                    # missing field `objects`
    (CAM + "[[objects]]\nbrdf = { type = \"lambert\", kd = [1.0,1.0,1.0] }\ngeometry = { type = \"sphere\", pos = [0.0,0.0,0.0], r = 1.0 }\n", "Parse"),
    (CAM + "[[objects]]\nbrdf = { type = \"diffuse\", kd = [1.0,1.0] }\ngeometry = { type = \"sphere\", pos = [0.0,0.0,0.0], r = 1.0 }\n", "Parse"),
    (CAM + "[[objects]]\nbrdf = { type = \"diffuse\", kd = [1.0,1.0,1.0] }\ngeometry = { type = \"torus\" }\n", "Parse"),
    (CAM + LIGHT + "transforms = [ { shear = 1.0 } ]\n", "Parse"),
    (CAM + "[[objects]]\nbrdf = { type = \"phong\", kd = 0.5, ks = 0.5, color_d = [1.0,1.0,1.0], color_s = [1.0,1.0,1.0], power = -2 }\ngeometry = { type = \"sphere\", pos = [0.0,0.0,0.0], r = 1.0 }\n" + LIGHT, "Parse"),
    (CAM + "[[objects]\n", "Parse"),
    (CAM + "x = \n", "Parse"),
    (CAM + "[[objects]]\nbrdf = { type = \"diffuse\", kd = [1.0,1.0,1.0] }\ngeometry = { type = \"mesh\", path = \"missing.obj\" }\n" + LIGHT, "MeshLoad"),
    (CAM + "[[objects]]\nbrdf = { type = \"diffuse\", kd = [1.0,1.0,1.0] }\ngeometry = { type = \"sphere\", pos = [0.0,0.0,0.0], r = 1.0 }\n", "NoLight"),
    (CAM + "[[objects]]\nemitted = [1.0,1.0,1.0]\nbrdf = { type = \"diffuse\", kd = [1.0,1.0,1.0] }\ngeometry = { type = \"plane\", pos = [0.0,0.0,0.0], n = [0.0,1.0,0.0] }\n", "Unsupported"),
])
def test_loader_errors(rtb, text, kind):
    with pytest.raises(rtb.LoadTomlError) as e:
        load(rtb, text)
    assert e.value.kind == kind


def test_loader_io_and_mesh_errors(rtb, tmp_path):
    with pytest.raises(rtb.LoadTomlError) as e:
        rtb.Scene.from_toml(str(tmp_path / "nope.toml"), device=-1)
    assert e.value.kind == "Io"
    (tmp_path / "bad.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 x\nf 1 2 3\n")
    (tmp_path / "range.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 9\n")
    (tmp_path / "short.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2\n")
    (tmp_path / "ok.obj").write_text("# c\nvn 0 0 1\nvt 0 0\ng grp\nv 0 0 0\nv 1 0 0\nv 0 1 0\nv 0 0 1\nf 1/7/1 2/8/1 3/9/1\nf 1//2 3//2 4//2 2//2\n")
    for bad in ("bad.obj", "range.obj", "short.obj"):
        with pytest.raises(rtb.LoadTomlError) as e:
            load(rtb, CAM + f"[[objects]]\nbrdf = {{ type = \"diffuse\", kd = [1.0,1.0,1.0] }}\ngeometry = {{ type = \"mesh\", path = \"{bad}\" }}\n" + LIGHT, assets=str(tmp_path))
        assert e.value.kind == "MeshLoad"
    sc = load(rtb, CAM + "[[objects]]\nbrdf = { type = \"diffuse\", kd = [1.0,1.0,1.0] }\ngeometry = { type = \"mesh\", path = \"ok.obj\" }\n" + LIGHT, assets=str(tmp_path))
    # a/b/c tokens keep the first index; a 4th vertex on a face line is ignored (triangles only)
    assert sc.triangles().tolist() == [[[0, 0, 0], [1, 0, 0], [0, 1, 0]], [[0, 0, 0], [0, 1, 0], [0, 0, 1]]]


def test_host_only_handles_refuse_to_compute(rtb):
    sc = rtb.Scene.from_toml(scene_path("cornell_box"), device=-1)
    with pytest.raises(rtb.RtbError) as e:
        sc.render(8, 8, 4)
    assert e.value.code == rtb.RTB_ECUDA if hasattr(rtb, "RTB_ECUDA") else True
    with pytest.raises(rtb.RtbError):
        sc.trace_primary(4, 4)


def test_programmatic_scene_matches_toml(rtb):
    # rtb_scene_create: the same objects handed over in memory instead of TOML
    toml_scene = load(rtb, CAM + """
[[objects]]
brdf = { type = "diffuse", kd = [0.5, 0.6, 0.7] }
geometry = { type = "prism", pos = [1.0, 2.0, 3.0], size = [1.0, 2.0, 4.0] }
[[objects]]
brdf = { type = "specular", ks = [0.9, 0.9, 0.9] }
geometry = { type = "plane", pos = [0.0, -1.0, 0.0], n = [0.0, 1.0, 0.0] }
""" + LIGHT)
    tris = toml_scene.triangles()
    sc = rtb.Scene.from_objects((0, 0, 10), (0, 0, -1), [
        {"brdf": ("diffuse", (0.5, 0.6, 0.7)), "geometry": ("mesh", tris)},
        {"brdf": ("specular", (0.9, 0.9, 0.9)), "geometry": ("plane", (0, -1, 0), (0, 1, 0))},
        {"emitted": (1, 1, 1), "brdf": ("diffuse", (0, 0, 0)), "geometry": ("sphere", (0, 9, 0), 1.0)},
    ], device=-1)
    assert sc.info.n_objects == 3 and sc.light_source == 2 and sc.info.n_triangles == 12
    assert np.array_equal(sc.triangles(), tris)
    for i in range(3):
        a, b = sc.object(i), toml_scene.object(i)
        for key in ("brdf", "geometry", "emitted", "k", "pos", "n", "r", "n_triangles"):
            assert a[key] == b[key], (i, key)
        assert a["surface_area"] == pytest.approx(b["surface_area"], rel=1e-6)
    with pytest.raises(rtb.LoadTomlError) as e:
        rtb.Scene.from_objects((0, 0, 10), (0, 0, -1), [{"brdf": ("diffuse", (1, 1, 1)), "geometry": ("sphere", (0, 0, 0), 1.0)}], device=-1)
    assert e.value.kind == "NoLight"


@pytest.mark.parametrize("name", ["cubes", "flying_unicorn"])
def test_reference_octree_build_matches_oracle(rtb, oracle_scene, name):
    # Octree::build (src/geometry.rs:1149-1216) restated twice — in the product (octree_host.cpp, what
    # ACCEL_OCTREE_REFERENCE traverses on the device) and in the oracle: same node / leaf / triangle-reference census.
    # flying_unicorn: SURVEY's probe of the reference structure (47 183 nodes, 187 766 references).
    sc = rtb.Scene.from_toml(scene_path(name), device=-1)
    osc = oracle_scene(name)
    seen = 0
    for i in range(sc.info.n_objects):
        st = osc.mesh_stats(i)
        if st is None:
            with pytest.raises(rtb.RtbError):
                sc.octree_stats(i)
            continue
        mine = sc.octree_stats(i)
        assert (mine["parents"], mine["leaves"], mine["tri_refs"]) == (st["octree_parents"], st["octree_leaves"], st["octree_tri_refs"])
        seen += 1
    assert seen == (2 if name == "cubes" else 1)
    if name == "flying_unicorn":
        assert mine == {"nodes": 47183, "parents": 9540, "leaves": 37643, "tri_refs": 187766}


def test_scene_export_import_round_trip(rtb):
    # rtb_scene_export / rtb_scene_import (the multi-GPU scene broadcast), host part: every object field, every triangle and
    # the reference-octree census survive the blob; damaged blobs are refused with RTB_EPARSE
    a = rtb.Scene.from_toml(scene_path("flying_unicorn"), device=-1)
    blob = a.export()
    b = rtb.Scene.from_export(blob, device=-1)
    assert b.info.n_objects == a.info.n_objects == 9 and b.info.n_triangles == a.info.n_triangles == 37380
    assert b.light_source == a.light_source and list(b.info.camera_dir) == list(a.info.camera_dir)
    for i in range(a.info.n_objects):
        assert a.object(i) == b.object(i)
    assert np.array_equal(a.triangles(), b.triangles())
    assert a.octree_stats(6) == b.octree_stats(6)
    for bad in (blob[:100], blob[:-8], np.concatenate([blob, np.zeros(4, np.uint8)])):
        with pytest.raises(rtb.LoadTomlError) as e:
            rtb.Scene.from_export(bad, device=-1)
        assert e.value.kind == "Parse"
    broken = blob.copy()
    broken[0] ^= 0xFF
    with pytest.raises(rtb.LoadTomlError):
        rtb.Scene.from_export(broken, device=-1)
