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
