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


def test_struct_layouts_match_header(rtb, tmp_path):
    # sizes and a few field offsets as a C compiler sees include/rtb200.h, against the ctypes mirrors
    import subprocess

    from raytracer_server_b200 import _abi

    src = tmp_path / "layout.c"
    src.write_text("""#include <stdio.h>
#include <stddef.h>
#include "rtb200.h"
int main(void) {
    printf("%zu %zu %zu %zu %zu %zu ", sizeof(rtb_params), sizeof(rtb_scene_info), sizeof(rtb_stats), sizeof(rtb_object_info),
           sizeof(rtb_object_desc), sizeof(rtb_scene_desc));
    printf("%zu %zu %zu %zu %d\\n", offsetof(rtb_params, seed), offsetof(rtb_scene_info, bvh_min), offsetof(rtb_stats, rays_bvh),
           offsetof(rtb_stats, wall_ms), RTB_ABI_VERSION);
    return 0;
}
""")
    exe = tmp_path / "layout"
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    v = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert v[:6] == [C.sizeof(_abi.Params), C.sizeof(_abi.SceneInfo), C.sizeof(_abi.Stats), C.sizeof(_abi.ObjectInfo),
                     C.sizeof(_abi.ObjectDesc), C.sizeof(_abi.SceneDesc)]
    assert v[6:10] == [_abi.Params.seed.offset, _abi.SceneInfo.bvh_min.offset, _abi.Stats.rays_bvh.offset, _abi.Stats.wall_ms.offset]
    assert v[10] == 2 and C.sizeof(_abi.Params) == 56


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
