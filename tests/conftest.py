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
