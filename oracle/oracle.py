"""ctypes wrapper around liboracle.so — the CPU f64 ORACLE (test infrastructure only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs
may import this module.  The scene TOML is parsed with Python's ``tomllib`` (a full TOML
parser, independent of the product's C++ TOML-subset parser) following
``SceneSpec``/``to_scene`` in the reference (src/scene.rs:292-441); geometry, transforms,
octree build and all rendering happen in rt_oracle.cpp.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tomllib

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

ACCEL_OCTREE_FAITHFUL = 0
ACCEL_EXACT = 1
EST_NEE = 0
EST_MIS_DEAD = 1


def build(force: bool = False) -> str:
    """Compile liboracle.so (and its fp32-storage twin) with the Makefile next to this file (g++ only)."""
    src = os.path.join(_HERE, "rt_oracle.cpp")
    for name in ("liboracle.so", "liboracle_f32.so"):
        path = os.path.join(_HERE, name)
        if force or not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            subprocess.run(["make", "-C", _HERE, "-B", name], check=True, capture_output=True)
    return _LIB_PATH


_libs = {}


def lib(f32: bool = False):
    """liboracle.so (f64, THE oracle) or liboracle_f32.so (same restatement with fp32 storage: a debugging aid for the
    intersection hooks only — see ORACLE_F32 in rt_oracle.cpp)."""
    if f32 not in _libs:
        path = os.path.join(_HERE, "liboracle_f32.so") if f32 else _LIB_PATH
        src = os.path.join(_HERE, "rt_oracle.cpp")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            subprocess.run(["make", "-C", _HERE, "-B", os.path.basename(path)], check=True, capture_output=True)
        L = C.CDLL(path)
        dp = C.POINTER(C.c_float if f32 else C.c_double)
        real = C.c_float if f32 else C.c_double
        L.or_scene_new.restype = C.c_void_p
        L.or_scene_new.argtypes = [dp, dp]
        L.or_scene_free.argtypes = [C.c_void_p]
        L.or_last_error.restype = C.c_char_p
        L.or_last_error.argtypes = [C.c_void_p]
        L.or_add_object.restype = C.c_int
        L.or_add_object.argtypes = [C.c_void_p, dp, C.c_int, dp, C.c_int, dp, C.c_char_p, C.c_int,
                                    C.POINTER(C.c_int), dp]
        L.or_scene_finish.restype = C.c_long
        L.or_scene_finish.argtypes = [C.c_void_p]
        L.or_set_modes.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.or_num_objects.restype = C.c_long
        L.or_num_objects.argtypes = [C.c_void_p]
        L.or_mesh_stats.restype = C.c_long
        L.or_mesh_stats.argtypes = [C.c_void_p, C.c_long, dp, C.POINTER(C.c_long)]
        L.or_mesh_triangles.restype = C.c_long
        L.or_mesh_triangles.argtypes = [C.c_void_p, C.c_long, dp, C.c_long]
        L.or_octants.argtypes = [dp, dp, dp]
        L.or_trace_rays.argtypes = [C.c_void_p, C.c_long, dp, dp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), dp,
                                    dp, dp, C.POINTER(C.c_long)]
        L.or_primary_rays.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, real, real, dp, dp]
        L.or_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.POINTER(C.c_uint8), dp, C.POINTER(C.c_long)]
        L.or_sample_radiance.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_long,
                                         C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), dp]
        L.or_philox.argtypes = [C.c_uint32] * 6 + [C.POINTER(C.c_uint32)]
        _libs[f32] = L
    return _libs[f32]


def _dp(a):
    """pointer to a float64 or float32 array (the two builds of the library take double* / float*)"""
    return a.ctypes.data_as(C.POINTER(C.c_float if a.dtype == np.float32 else C.c_double))


def _vec64(v, n=3):
    a = np.asarray(v, dtype=np.float64).reshape(-1)
    if a.size != n:
        raise ValueError(f"expected {n} numbers, got {a.size}")
    return np.ascontiguousarray(a)


class OracleError(RuntimeError):
    pass


class OracleScene:
    """Scene::from_toml (src/scene.rs:143-150) + SceneSpec::to_scene (src/scene.rs:357-441)."""

    def __init__(self, spec: dict, assets_dir: str | None = None, f32: bool = False):
        """f32 = True: the fp32-storage build (liboracle_f32.so) — only primary_rays / trace_rays are meaningful there."""
        self._f32 = f32
        self._dt = np.float32 if f32 else np.float64
        L = self._L = lib(f32)
        _vec = lambda v, n=3: _vec64(v, n).astype(self._dt)   # noqa: E731
        cam = spec["camera"]
        pos, d = _vec(cam["pos"]), _vec(cam["dir"])
        self._h = L.or_scene_new(_dp(pos), _dp(d))
        self.spec = spec
        for ob in spec["objects"]:
            emitted = _vec(ob.get("emitted", [0.0, 0.0, 0.0]))
            b = ob["brdf"]
            bt = b["type"]
            if bt == "diffuse":
                bk, bp = 0, _vec(b["kd"])
            elif bt == "specular":
                bk, bp = 1, _vec(b["ks"])
            elif bt == "phong":
                if int(b["power"]) < 0:
                    raise OracleError("parse: power must be usize")
                bk = 2
                bp = np.array([float(b["kd"]), float(b["ks"]), float(int(b["power"]))]
                              + list(map(float, b["color_d"])) + list(map(float, b["color_s"])), dtype=self._dt)
            else:
                raise OracleError(f"parse: unknown brdf type {bt}")
            g = ob["geometry"]
            gt = g["type"]
            path = b""
            if gt == "sphere":
                gk, gp = 0, np.array(list(map(float, g["pos"])) + [float(g["r"])])
            elif gt == "plane":
                gk, gp = 1, np.array(list(map(float, g["pos"])) + list(map(float, g["n"])))
            elif gt == "mesh":
                gk, gp = 2, np.zeros(6)
                if assets_dir is None:
                    raise OracleError("mesh geometry needs assets_dir")
                path = os.path.join(assets_dir, g["path"]).encode()
            elif gt == "cube":
                s = float(g["size"])
                gk, gp = 3, np.array(list(map(float, g["pos"])) + [s, s, s])
            elif gt == "prism":
                gk, gp = 3, np.array(list(map(float, g["pos"])) + list(map(float, g["size"])))
            else:
                raise OracleError(f"parse: unknown geometry type {gt}")
            gp = np.ascontiguousarray(gp, dtype=self._dt)
            kinds, vals = [], []
            for t in ob.get("transforms", []) or []:
                (k, v), = t.items()
                if k == "translate":
                    kinds.append(0); vals.append(list(map(float, v)))
                elif k == "scale":
                    kinds.append(1); vals.append([float(v), 0.0, 0.0])
                elif k in ("rotate_x", "rotate_y", "rotate_z"):
                    kinds.append({"rotate_x": 2, "rotate_y": 3, "rotate_z": 4}[k]); vals.append([float(v), 0.0, 0.0])
                else:
                    raise OracleError(f"parse: unknown transform {k}")
            ka = (C.c_int * max(1, len(kinds)))(*kinds)
            va = np.ascontiguousarray(np.array(vals, dtype=self._dt).reshape(-1)) if vals else np.zeros(3, dtype=self._dt)
            rc = L.or_add_object(self._h, _dp(emitted), bk, _dp(np.ascontiguousarray(bp)), gk, _dp(gp), path,
                                 len(kinds), ka, _dp(va))
            if rc != 0:
                raise OracleError(L.or_last_error(self._h).decode())
        self.light_source = L.or_scene_finish(self._h)
        if self.light_source < 0:
            raise OracleError("no light source (reference: unreachable!() src/scene.rs:136)")
        self.set_modes(ACCEL_EXACT, EST_NEE)

    @classmethod
    def from_toml(cls, path: str, assets_dir: str | None = None, f32: bool = False) -> "OracleScene":
        with open(path, "rb") as f:
            spec = tomllib.load(f)
        if assets_dir is None:
            assets_dir = os.path.join(os.path.dirname(os.path.abspath(path)), "assets")
        return cls(spec, assets_dir, f32)

    @classmethod
    def from_toml_string(cls, text: str, assets_dir: str | None = None, f32: bool = False) -> "OracleScene":
        return cls(tomllib.loads(text), assets_dir, f32)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._L.or_scene_free(self._h)
                self._h = None
        except Exception:
            pass

    def set_modes(self, accel: int = ACCEL_EXACT, estimator: int = EST_NEE):
        self.accel, self.estimator = accel, estimator
        self._L.or_set_modes(self._h, accel, estimator)

    @property
    def num_objects(self) -> int:
        return self._L.or_num_objects(self._h)

    def mesh_stats(self, obj: int):
        bbox = np.zeros(6, dtype=self._dt)
        cnt = (C.c_long * 4)()
        if self._L.or_mesh_stats(self._h, obj, _dp(bbox), cnt) != 0:
            return None
        return {"bbox_min": bbox[:3].copy(), "bbox_max": bbox[3:].copy(), "triangles": cnt[0],
                "octree_parents": cnt[1], "octree_leaves": cnt[2], "octree_tri_refs": cnt[3]}

    def mesh_triangles(self, obj: int) -> np.ndarray:
        n = self._L.or_mesh_triangles(self._h, obj, _dp(np.zeros(9, dtype=self._dt)), 0)
        if n < 0:
            return np.zeros((0, 3, 3))
        out = np.zeros((n, 3, 3), dtype=self._dt)
        self._L.or_mesh_triangles(self._h, obj, _dp(out), n)
        return out

    def trace_rays(self, org, dirs, want_geom: bool = False, count_work: bool = False):
        org = np.ascontiguousarray(org, dtype=self._dt).reshape(-1, 3)
        dirs = np.ascontiguousarray(dirs, dtype=self._dt).reshape(-1, 3)
        n = org.shape[0]
        obj = np.empty(n, dtype=np.int32)
        tri = np.empty(n, dtype=np.int32)
        t = np.empty(n, dtype=self._dt)
        pos = np.empty((n, 3), dtype=self._dt) if want_geom else None
        nrm = np.empty((n, 3), dtype=self._dt) if want_geom else None
        work = (C.c_long * 3)(0, 0, 0)
        self._L.or_trace_rays(self._h, n, _dp(org), _dp(dirs), obj.ctypes.data_as(C.POINTER(C.c_int32)),
                            tri.ctypes.data_as(C.POINTER(C.c_int32)), _dp(t),
                            _dp(pos) if want_geom else None, _dp(nrm) if want_geom else None,
                            work if count_work else None)
        res = {"obj": obj, "tri": tri, "t": t}
        if want_geom:
            res["pos"], res["n"] = pos, nrm
        if count_work:
            res["work"] = {"node_visits": work[0], "box_tests": work[1], "tri_tests": work[2]}
        return res

    def primary_rays(self, width: int, height: int, sx: int = 0, sy: int = 0, dx: float = 0.0, dy: float = 0.0):
        org = np.empty((height * width, 3), dtype=self._dt)
        dirs = np.empty((height * width, 3), dtype=self._dt)
        self._L.or_primary_rays(self._h, width, height, sx, sy, dx, dy, _dp(org), _dp(dirs))
        return org, dirs

    def render(self, width: int, height: int, spp: int, seed: int = 0, y0: int = 0, y1: int | None = None,
               nthreads: int = 1, want_sub: bool = False, row_stride: int = 1):
        """Returns dict(rgb8 [h,w,3] uint8, sub [h,w,4,3] f64 or None, rays, samples).
        nthreads > 0: static row bands like src/server.rs:166-168; < 0: |n| threads, dynamic rows.
        Only rows y0, y0+row_stride, ... < y1 are rendered."""
        if y1 is None:
            y1 = height
        rgb = np.zeros((height, width, 3), dtype=np.uint8)
        sub = np.zeros((height, width, 4, 3), dtype=self._dt) if want_sub else None
        cnt = (C.c_long * 2)(0, 0)
        self._L.or_render(self._h, width, height, spp, seed, y0, y1, row_stride, nthreads,
                        rgb.ctypes.data_as(C.POINTER(C.c_uint8)), _dp(sub) if want_sub else None, cnt)
        return {"rgb8": rgb, "sub": sub, "rays": cnt[0], "samples": cnt[1]}

    def sample_radiance(self, width, height, spp, seed, px, py_img, sample_idx):
        px = np.ascontiguousarray(px, dtype=np.int32)
        py = np.ascontiguousarray(py_img, dtype=np.int32)
        si = np.ascontiguousarray(sample_idx, dtype=np.int32)
        out = np.empty((px.size, 3), dtype=self._dt)
        ip = C.POINTER(C.c_int32)
        self._L.or_sample_radiance(self._h, width, height, spp, seed, px.size, px.ctypes.data_as(ip),
                                 py.ctypes.data_as(ip), si.ctypes.data_as(ip), _dp(out))
        return out


def octants(mn, mx) -> np.ndarray:
    out = np.zeros((8, 2, 3))
    lib().or_octants(_dp(_vec64(mn)), _dp(_vec64(mx)), _dp(out))
    return out


def philox4x32_10(counter, key):
    out = (C.c_uint32 * 4)()
    lib().or_philox(*[int(c) & 0xFFFFFFFF for c in counter], *[int(k) & 0xFFFFFFFF for k in key], out)
    return [int(x) for x in out]
