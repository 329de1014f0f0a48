"""The C-ABI library loads without a GPU and exports every symbol include/rtb200.h declares;
the ctypes mirrors have the header's struct sizes."""
import ctypes as C
import os
import re

from conftest import ROOT


def header_functions():
    text = open(os.path.join(ROOT, "include", "rtb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rtb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(rtb):
    from raytracer_server_b200 import _abi

    L = _abi.lib()
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in rtb200.h but not exported by librtb200.so"
    assert sorted(_abi.EXPORTS) == names


def test_struct_layouts_match_header(rtb):
    from raytracer_server_b200 import _abi

    assert C.sizeof(_abi.Params) == 4 * 4 + 8 + 3 * 4 + 5 * 4 + 0 or C.sizeof(_abi.Params) == 56
    assert C.sizeof(_abi.Params) == 56
    assert C.sizeof(_abi.SceneInfo) == 12 * 4 + 12 * 4 + 8
    assert C.sizeof(_abi.Stats) == 8 * 8 + 6 * 8 + 3 * 8
    assert C.sizeof(_abi.ObjectInfo) == 4 * 4 + 8 * (3 * 6 + 1 + 3 * 2 + 1)


def test_param_validation_is_cpu_side(rtb):
    from raytracer_server_b200 import _abi

    p = rtb.make_params(100, 70, 8, rank=1, world=4)
    assert _abi.lib().rtb_local_pixels(C.byref(p)) == 3 * 1024          # 4 x 3 tiles, ranks 0..3 own 3 each
    bad = rtb.make_params(0, 70, 8)
    assert _abi.lib().rtb_local_pixels(C.byref(bad)) == _abi.RTB_EINVAL
    bad = rtb.make_params(10, 10, 8, rank=2, world=2)
    assert _abi.lib().rtb_local_pixels(C.byref(bad)) == _abi.RTB_EINVAL
    assert "rank" in _abi.last_error()


def test_no_cpu_fallback_without_device(rtb):
    import torch

    if torch.cuda.is_available():
        return
    import pytest

    from conftest import scene_path

    with pytest.raises(rtb.RtbError) as e:
        rtb.Scene.from_toml(scene_path("cornell_box"), device=0)
    assert e.value.code == _abi_code(rtb)


def _abi_code(rtb):
    from raytracer_server_b200 import _abi

    return _abi.RTB_ECUDA
