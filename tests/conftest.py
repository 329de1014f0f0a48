import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
SCENES = os.path.join(ROOT, "tests", "golden", "scenes")
SCENE_NAMES = ("cornell_box", "cubes", "flying_unicorn")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def scene_path(name):
    return os.path.join(SCENES, name + ".toml")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as O

    O.build()
    return O


_oracle_cache = {}


@pytest.fixture(scope="session")
def oracle_scene(oracle_mod):
    def get(name):
        if name not in _oracle_cache:
            _oracle_cache[name] = oracle_mod.OracleScene.from_toml(scene_path(name))
        sc = _oracle_cache[name]
        sc.set_modes(oracle_mod.ACCEL_EXACT, oracle_mod.EST_NEE)
        return sc

    return get


@pytest.fixture(scope="session")
def rtb():
    import raytracer_server_b200 as R

    return R


_gpu_cache = {}


@pytest.fixture(scope="session")
def gpu_scene(rtb):
    def get(name):
        if name not in _gpu_cache:
            _gpu_cache[name] = rtb.Scene.from_toml(scene_path(name), device=0)
        return _gpu_cache[name]

    return get


NCPU = os.cpu_count() or 1


@pytest.fixture(scope="session")
def oracle_f32_twin(oracle_mod, gpu_scene, tmp_path_factory):
    """The fp32-storage build of the oracle (liboracle_f32.so) fed with EXACTLY the geometry the GPU holds: analytic objects
    from the loader's records (rounded to fp32 on both sides), meshes from the device scene's own fp32 triangles (written to
    an OBJ with exact decimal representations, no transforms).  Answers "would the reference's arithmetic, on fp32 data,
    decide like the GPU?" for rays on which the GPU and the f64 oracle disagree."""
    cache = {}

    def get(name):
        if name in cache:
            return cache[name]
        g = gpu_scene(name)
        d = tmp_path_factory.mktemp("f32twin_" + name)
        tris = g.triangles()
        info = g.info
        objects = []
        for i in range(info.n_objects):
            ob = g.object(i)
            brdf = ({"type": "diffuse", "kd": ob["k"]} if ob["brdf"] == 0 else {"type": "specular", "ks": ob["k"]} if ob["brdf"] == 1 else
                    {"type": "phong", "kd": ob["k"][0], "ks": ob["k"][1], "power": int(ob["k"][2]), "color_d": ob["color_d"], "color_s": ob["color_s"]})
            if ob["geometry"] == 0:
                geom = {"type": "sphere", "pos": ob["pos"], "r": ob["r"]}
            elif ob["geometry"] == 1:
                geom = {"type": "plane", "pos": ob["pos"], "n": ob["n"]}
            else:
                t = tris[ob["first_triangle"]: ob["first_triangle"] + ob["n_triangles"]].reshape(-1, 3)
                with open(d / f"mesh{i}.obj", "w") as f:
                    for v in t:
                        f.write(f"v {float(v[0])!r} {float(v[1])!r} {float(v[2])!r}\n")
                    for k in range(ob["n_triangles"]):
                        f.write(f"f {3 * k + 1} {3 * k + 2} {3 * k + 3}\n")
                geom = {"type": "mesh", "path": f"mesh{i}.obj"}
            objects.append({"emitted": ob["emitted"], "brdf": brdf, "geometry": geom})
        spec = {"camera": {"pos": list(info.camera_pos), "dir": list(info.camera_dir)}, "objects": objects}
        cache[name] = oracle_mod.OracleScene(spec, str(d), f32=True)
        cache[name].set_modes(oracle_mod.ACCEL_EXACT, oracle_mod.EST_NEE)
        return cache[name]

    return get


# ---- measured parity numbers: every parity test records what it measured (not only whether it passed); the session
# writes them to gpurun_out/parity_{gpu,cpu}.json (override: $RTB_PARITY_OUT) — the GPU file is committed per round as
# profiles/parity_rNN.json
_parity = {}


@pytest.fixture(scope="session")
def parity_log():
    def log(key, **values):
        def conv(v):
            import numpy as np

            if isinstance(v, np.ndarray):
                return [conv(x) for x in v.tolist()]
            if isinstance(v, (np.floating, np.integer, np.bool_)):
                return v.item()
            if isinstance(v, (list, tuple)):
                return [conv(x) for x in v]
            return v

        _parity.setdefault(key, {}).update({k: conv(v) for k, v in values.items()})

    return log


def pytest_sessionfinish(session, exitstatus):
    if not _parity:
        return
    import json

    gpu = any(k.startswith("gpu/") for k in _parity)
    path = os.environ.get("RTB_PARITY_OUT") or os.path.join(ROOT, "gpurun_out", "parity_gpu.json" if gpu else "parity_cpu.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        old = {}
        if os.path.exists(path):
            with open(path) as f:
                old = json.load(f)
        old.update(_parity)
        with open(path, "w") as f:
            json.dump(old, f, indent=1, sort_keys=True)
    except Exception as e:  # never fail a test run over the report
        print(f"parity log not written: {e}")
