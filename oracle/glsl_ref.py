"""ctypes binding of oracle/_ref/libvct_glsl_ref.so: THE REFERENCE'S OWN GLSL executed on the CPU.

TEST INFRASTRUCTURE ONLY.  The library is built by `make -C oracle ref` where /root/reference exists (the shader text and GLM are
read from where they lie, see oracle/glsl_ref/harness.cpp); it is git-ignored and travels to the GPU box prebuilt.  Two sets of
entry points: mode "rules" (built-ins evaluated by the oracle's written rules R5 / R9 -- must equal the oracle bit for bit) and
mode "glm" (GLM's own built-ins).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import orc

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libvct_glsl_ref.so")
REFERENCE = "/root/reference"
_LIB = None
MODES = ("rules", "glm")


def available() -> bool:
    """True when the library exists or can be built here (needs the reference tree)."""
    return os.path.exists(SO) or os.path.isfile(os.path.join(REFERENCE, "shader", "voxelize.frag"))


def lib():
    global _LIB
    if _LIB is None:
        if os.path.isfile(os.path.join(REFERENCE, "shader", "voxelize.frag")):
            subprocess.check_call(["make", "-C", _HERE, "_ref/libvct_glsl_ref.so"], stdout=subprocess.DEVNULL)
        L = C.CDLL(SO)
        for m in MODES:
            f = lambda n: getattr(L, "glref_%s_%s" % (m, n))  # noqa: E731
            f("voxelize").argtypes = [C.POINTER(orc.SceneT), C.c_int, C.c_void_p, C.POINTER(C.c_uint64)]
            f("voxelize_variant").argtypes = [C.POINTER(orc.SceneT), C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.c_int, C.c_int, C.c_uint32]
            f("mipmap").argtypes = [C.c_void_p, C.c_int, C.c_int]
            f("gbuffer").argtypes = [C.POINTER(orc.SceneT), orc.f32p, orc.f32p, C.c_int, C.c_int] + [C.c_void_p] * 5
            f("shade").argtypes = [C.POINTER(orc.SceneT), orc.f32p, C.c_int, C.c_int] + [C.c_void_p] * 4 + [
                C.c_void_p, C.c_int, C.c_int, C.POINTER(orc.TraceParams), C.c_int, C.c_int, C.c_void_p]
            f("trace_cone").argtypes = [C.c_void_p, C.c_int, C.c_int, orc.f32p, orc.f32p, C.c_float, C.c_float, orc.f32p]
            f("rgba8_avg").restype = C.c_uint32
            f("rgba8_avg").argtypes = [C.c_uint32, orc.f32p]
            f("select_axis").argtypes = [orc.f32p, orc.f32p, orc.f32p]
        L.glref_camera.argtypes = [orc.f32p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, orc.f32p, orc.f32p]
        L.glref_model_trs.argtypes = [orc.f32p, C.c_float, C.c_float, orc.f32p]
        _LIB = L
    return _LIB


def _fn(mode: str, name: str):
    assert mode in MODES
    return getattr(lib(), "glref_%s_%s" % (mode, name))


def voxelize(scene, R: int, mode: str = "rules"):
    """-> (six level-0 images [6][R][R][R] uint32, fragments executed)"""
    sr = orc.SceneRef(scene)
    tex = np.zeros((6, R, R, R), np.uint32)
    ptrs = (C.c_void_p * 6)(*[tex[i].ctypes.data for i in range(6)])
    n = C.c_uint64(0)
    rc = _fn(mode, "voxelize")(C.byref(sr.c), R, ptrs, C.byref(n))
    assert rc == 0, rc
    return tex, int(n.value)


ORDER_RULE, ORDER_REVERSED_IN_TRIANGLE, ORDER_REVERSED_TRIANGLES, ORDER_RANDOM = 0, 1, 2, 3
BARY_SNAPPED, BARY_UNSNAPPED = 0, 1


def voxelize_variant(scene, R: int, order: int = ORDER_RULE, bary: int = BARY_SNAPPED, seed: int = 0, mode: str = "rules") -> np.ndarray:
    """SENSITIVITY STUDIES ONLY (tools/fixed_function_sensitivity.py): the voxel grid under another valid fragment order (GL guarantees none)
    or with barycentrics from the unsnapped vertex positions; -> texture 0"""
    sr = orc.SceneRef(scene)
    tex = np.zeros((6, R, R, R), np.uint32)
    ptrs = (C.c_void_p * 6)(*[tex[i].ctypes.data for i in range(6)])
    n = C.c_uint64(0)
    rc = _fn(mode, "voxelize_variant")(C.byref(sr.c), R, ptrs, C.byref(n), order, bary, seed)
    assert rc == 0, rc
    return tex[0]


def mipmap(base: np.ndarray, n_levels: int = 7, mode: str = "rules") -> orc.Pyramid:
    p = orc.Pyramid(base, n_levels)
    rc = _fn(mode, "mipmap")(p.ptrs, p.R, n_levels)
    assert rc == 0, rc
    return p


def gbuffer(scene, view, proj, W: int, H: int, mode: str = "rules") -> orc.GBuffer:
    sr = orc.SceneRef(scene)
    g = orc.GBuffer(W, H)
    rc = _fn(mode, "gbuffer")(C.byref(sr.c), orc._fp(view), orc._fp(proj), W, H, g.tri_id.ctypes.data, g.depth.ctypes.data,
                              g.world_pos.ctypes.data, g.normal.ctypes.data, g.material.ctypes.data)
    assert rc == 0, rc
    return g


def shade(scene, view, g: orc.GBuffer, p: orc.Pyramid, params=None, tile_stride: int = 1, tile_phase: int = 0, mode: str = "rules"):
    assert p.fmt == orc.FMT_RGBA8
    sr = orc.SceneRef(scene)
    params = params or orc.default_params()
    assert params.n_diffuse_cones == 9, "the reference shader traces 9 diffuse cones"
    frame = np.zeros((g.H, g.W), np.uint32)
    rc = _fn(mode, "shade")(C.byref(sr.c), orc._fp(view), g.W, g.H, g.tri_id.ctypes.data, g.world_pos.ctypes.data, g.normal.ctypes.data,
                            g.material.ctypes.data, p.ptrs, p.R, p.n_levels, C.byref(params), tile_stride, tile_phase, frame.ctypes.data)
    assert rc == 0, rc
    return frame


def trace_cone(p: orc.Pyramid, origin, direction, aperture: float, max_dist: float, mode: str = "rules"):
    out = np.zeros(4, np.float32)
    rc = _fn(mode, "trace_cone")(p.ptrs, p.R, p.n_levels, orc._fp(origin), orc._fp(direction), aperture, max_dist, out.ctypes.data_as(orc.f32p))
    assert rc == 0, rc
    return out


def fold(stored: int, val01, mode: str = "rules") -> int:
    return int(_fn(mode, "rgba8_avg")(stored, orc._fp(val01)))


def select_axis(a, b, c, mode: str = "rules") -> int:
    return int(_fn(mode, "select_axis")(orc._fp(a), orc._fp(b), orc._fp(c)))


def render_frame(scene, view, proj, R: int, W: int, H: int, params=None, n_levels: int = 7, mode: str = "rules"):
    """Renderer::render() (src/renderer.cpp:392-405) with every programmable stage executed from the reference's GLSL"""
    tex, n_frag = voxelize(scene, R, mode)
    assert all(np.array_equal(tex[i], tex[0]) for i in range(1, 6)), "voxelize.frag:159-160 writes the same value to all six textures"
    pyr = mipmap(tex[0], n_levels, mode)
    g = gbuffer(scene, view, proj, W, H, mode)
    frame = shade(scene, view, g, pyr, params, mode=mode)
    return dict(base=tex[0], pyramid=pyr, gbuffer=g, frame=frame, fragments=n_frag)


def camera(eye, pitch_deg: float, yaw_deg: float, lens_angle: float, aspect: float, z_near: float, z_far: float):
    """the reference's own Camera struct (src/camera.h, compiled where it lies): -> (view, projection), column-major"""
    view, proj = np.zeros(16, np.float32), np.zeros(16, np.float32)
    rc = lib().glref_camera(orc._fp(eye), pitch_deg, yaw_deg, lens_angle, aspect, z_near, z_far, view.ctypes.data_as(orc.f32p), proj.ctypes.data_as(orc.f32p))
    assert rc == 0
    return view, proj


def model_trs(t, rot_y: float, scale: float) -> np.ndarray:
    """glm::translate / rotate / scale as src/main.cpp:369-372 applies them (identity start), column-major"""
    out = np.zeros(16, np.float32)
    rc = lib().glref_model_trs(orc._fp(t), rot_y, scale, out.ctypes.data_as(orc.f32p))
    assert rc == 0
    return out
