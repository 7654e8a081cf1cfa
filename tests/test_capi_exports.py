"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol declared in
include/vct/vct_c.h and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

from voxel_cone_tracing_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "vct", "vct_c.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vct_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = capi.load()
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), f"{n} declared in vct_c.h but not exported"
    assert sorted(capi.EXPORTS) == names


def test_struct_layouts_match_reference_sizes():
    from voxel_cone_tracing_b200 import scene as S
    assert S.VERTEX.itemsize == 32      # vert_data_t, renderer.cpp:24-35
    assert S.MATERIAL.itemsize == 128   # material_data_t, renderer.h:94-120
    assert S.LIGHT.itemsize == 28       # point_light_t, renderer.h:87-92
    assert S.MATERIAL.fields["emission"][1] == 64 and S.MATERIAL.fields["shininess"][1] == 76
    assert S.MATERIAL.fields["illum"][1] == 88 and S.MATERIAL.fields["anisotropy_rotation"][1] == 116
    assert ctypes.sizeof(capi.TraceParams) == 40


def test_no_cpu_fallback(gpu_available):
    if gpu_available:
        return
    L = capi.load()
    h = ctypes.c_void_p()
    assert L.vct_device_create(0, ctypes.byref(h)) == -2
    assert b"no CPU fallback" in L.vct_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "voxel_cone_tracing_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dp, f), errors="replace").read()
                assert "vct_oracle" not in src and "from oracle" not in src and "import oracle" not in src, f
