"""ctypes binding of the CPU oracle (oracle/libvct_oracle.so).

TEST INFRASTRUCTURE ONLY: the product (voxel_cone_tracing_b200) never imports this module.
PARITY: shader arithmetic pinned to the reference's own GLSL run on the CPU (oracle/glsl_ref.py); fixed-function GL rules held
against Mesa llvmpipe running the reference's passes (oracle/gl_ref.py); fragment order (R4) a written rule -- see oracle/vct_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

u32p = C.POINTER(C.c_uint32)
f32p = C.POINTER(C.c_float)


class SceneT(C.Structure):
    _fields_ = [("verts", C.c_void_p), ("n_verts", C.c_uint32), ("indices", C.c_void_p), ("n_indices", C.c_uint32),
                ("draws", C.c_void_p), ("n_draws", C.c_uint32), ("mats", C.c_void_p), ("n_mats", C.c_uint32),
                ("lights", C.c_void_p), ("n_lights", C.c_uint32), ("cube_size", C.c_float)]


class VoxelStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("fragments", "fragments_oob", "occupied", "max_per_voxel", "wrapped_voxels", "tris_no_frag")]


class TraceStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("shaded_pixels", "samples_diffuse", "samples_shadow", "samples_specular", "samples_refraction")]

    @property
    def samples(self):
        return self.samples_diffuse + self.samples_shadow + self.samples_specular + self.samples_refraction


class TraceParams(C.Structure):
    _fields_ = [("enable_direct", C.c_int), ("enable_diffuse", C.c_int), ("enable_specular", C.c_int), ("enable_shadow", C.c_int),
                ("view_voxel_dir", C.c_int), ("view_voxel_lod", C.c_float), ("n_diffuse_cones", C.c_int)]


def default_params(**kw) -> TraceParams:
    p = TraceParams(1, 1, 1, 1, 7, 0.0, 9)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libvct_oracle.so")
    src = [os.path.join(_HERE, f) for f in ("vct_oracle.cpp", "vct_oracle.h")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "libvct_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        try:
            L = C.CDLL(so)
        except OSError:
            L = C.CDLL(build(force=True))
        L.orc_rgba8_avg_fold.restype = C.c_uint32
        L.orc_rgba8_avg_fold.argtypes = [C.c_uint32, f32p]
        L.orc_select_axis.argtypes = [f32p, f32p, f32p]
        L.orc_perspective.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, f32p]
        L.orc_look_at.argtypes = [f32p, f32p, f32p, f32p]
        L.orc_camera_front.argtypes = [C.c_float, C.c_float, f32p]
        L.orc_specular_aperture.restype = C.c_float
        L.orc_specular_aperture.argtypes = [C.c_float]
        L.orc_trace_cone.argtypes = [C.c_void_p, C.c_int, C.c_int, f32p, f32p, C.c_float, C.c_float, f32p]
        L.orc_texture_lod.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, f32p, C.c_float, f32p]
        L.orc_voxelize.argtypes = [C.POINTER(SceneT), C.c_int, C.c_void_p, C.POINTER(VoxelStats)]
        L.orc_voxelize_slab.argtypes = [C.POINTER(SceneT), C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(VoxelStats)]
        L.orc_voxelize_slab_mode.argtypes = [C.POINTER(SceneT), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(VoxelStats)]
        L.orc_mipmap.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_mipmap_fmt.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_trace_cone_fmt.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, f32p, f32p, C.c_float, C.c_float, f32p]
        L.orc_trace_fmt.argtypes = [C.POINTER(SceneT), f32p, C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.c_void_p, C.c_int, C.c_int, C.c_int,
                                    C.POINTER(TraceParams), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(TraceStats)]
        L.orc_gbuffer.argtypes = [C.POINTER(SceneT), f32p, f32p, C.c_int, C.c_int] + [C.c_void_p] * 5
        L.orc_trace.argtypes = [C.POINTER(SceneT), f32p, C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.c_void_p, C.c_int, C.c_int,
                                C.POINTER(TraceParams), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(TraceStats)]
        L.orc_render_forward.argtypes = [C.POINTER(SceneT), f32p, f32p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(TraceParams), C.c_void_p]
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_debug_set_lod_filter.argtypes = [C.c_int]
        L.orc_debug_set_unorm_unpack.argtypes = [C.c_int]
        L.orc_debug_set_mip_balanced_sum.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def _fp(a):
    return np.ascontiguousarray(a, np.float32).ctypes.data_as(f32p)


def fold(stored: int, val01) -> int:
    return int(lib().orc_rgba8_avg_fold(stored, _fp(val01)))


def select_axis(a, b, c) -> int:
    return int(lib().orc_select_axis(_fp(a), _fp(b), _fp(c)))


def perspective(fovy, aspect, zn, zf):
    out = np.zeros(16, np.float32); lib().orc_perspective(fovy, aspect, zn, zf, out.ctypes.data_as(f32p)); return out


def look_at(eye, center, up):
    out = np.zeros(16, np.float32); lib().orc_look_at(_fp(eye), _fp(center), _fp(up), out.ctypes.data_as(f32p)); return out


def camera_front(pitch, yaw):
    out = np.zeros(3, np.float32); lib().orc_camera_front(pitch, yaw, out.ctypes.data_as(f32p)); return out


def specular_aperture(ns: float) -> float:
    return float(lib().orc_specular_aperture(ns))


class SceneRef:
    """Keeps the numpy buffers alive next to the C struct."""

    def __init__(self, scene):
        self.keep = [np.ascontiguousarray(scene.verts), np.ascontiguousarray(scene.indices), np.ascontiguousarray(scene.draws),
                     np.ascontiguousarray(scene.materials), np.ascontiguousarray(scene.lights)]
        v, i, d, m, l = self.keep
        self.c = SceneT(v.ctypes.data, len(v), i.ctypes.data, len(i), d.ctypes.data, len(d), m.ctypes.data, len(m),
                        l.ctypes.data if len(l) else None, len(l), float(scene.cube_size))


ACCUM_ORDERED, ACCUM_FIXED_POINT, ACCUM_FP16 = 0, 1, 2   # 2: fixed-point accumulation, RGBA16F voxels (uint64 = four halves)
FMT_RGBA8, FMT_RGBA16F = 0, 1


def voxelize(scene, R: int, z0: int = 0, z1: int | None = None, accum_mode: int = ACCUM_ORDERED):
    sr = SceneRef(scene)
    base = np.zeros((R, R, R), np.uint64 if accum_mode == ACCUM_FP16 else np.uint32)
    st = VoxelStats()
    rc = lib().orc_voxelize_slab_mode(C.byref(sr.c), R, z0, R if z1 is None else z1, accum_mode, base.ctypes.data, C.byref(st))
    assert rc == 0
    return base, st


class Pyramid:
    """Reference layout: 6 textures x n_levels (level 0 of all six aliases base)."""

    def __init__(self, base: np.ndarray, n_levels: int = 7, fmt: int = FMT_RGBA8):
        R = base.shape[0]
        self.R, self.n_levels, self.fmt = R, n_levels, fmt
        dt = np.uint64 if fmt == FMT_RGBA16F else np.uint32
        self.base = np.ascontiguousarray(base, dt)
        self.levels = [[self.base] + [np.zeros((max(R >> l, 1),) * 3, dt) for l in range(1, n_levels)] for _ in range(6)]
        self.ptrs = (C.c_void_p * (6 * n_levels))(*[self.levels[d][l].ctypes.data for d in range(6) for l in range(n_levels)])


def mipmap(base: np.ndarray, n_levels: int = 7, fmt: int = FMT_RGBA8) -> Pyramid:
    p = Pyramid(base, n_levels, fmt)
    rc = lib().orc_mipmap_fmt(p.base.ctypes.data, p.R, n_levels, p.ptrs, fmt)
    assert rc == 0
    return p


def texture_lod(p: Pyramid, d: int, pos, lod: float):
    out = np.zeros(4, np.float32); lib().orc_texture_lod(p.ptrs, p.R, p.n_levels, d, _fp(pos), lod, out.ctypes.data_as(f32p)); return out


def trace_cone(p: Pyramid, origin, direction, aperture: float, max_dist: float):
    out = np.zeros(4, np.float32)
    n = lib().orc_trace_cone_fmt(p.ptrs, p.R, p.n_levels, p.fmt, _fp(origin), _fp(direction), aperture, max_dist, out.ctypes.data_as(f32p))
    return out, int(n)


class GBuffer:
    def __init__(self, W, H):
        self.W, self.H = W, H
        self.tri_id = np.zeros((H, W), np.uint32); self.depth = np.zeros((H, W), np.float32)
        self.world_pos = np.zeros((H, W, 3), np.float32); self.normal = np.zeros((H, W, 3), np.float32)
        self.material = np.zeros((H, W), np.uint32)


def gbuffer(scene, view, proj, W: int, H: int) -> GBuffer:
    sr = SceneRef(scene); g = GBuffer(W, H)
    rc = lib().orc_gbuffer(C.byref(sr.c), _fp(view), _fp(proj), W, H, g.tri_id.ctypes.data, g.depth.ctypes.data,
                           g.world_pos.ctypes.data, g.normal.ctypes.data, g.material.ctypes.data)
    assert rc == 0
    return g


def trace(scene, view, g: GBuffer, p: Pyramid, params: TraceParams | None = None, tile_stride: int = 1, tile_phase: int = 0, frame=None):
    sr = SceneRef(scene)
    params = params or default_params()
    if frame is None:
        frame = np.zeros((g.H, g.W), np.uint32)
    st = TraceStats()
    rc = lib().orc_trace_fmt(C.byref(sr.c), _fp(view), g.W, g.H, g.tri_id.ctypes.data, g.world_pos.ctypes.data, g.normal.ctypes.data,
                             g.material.ctypes.data, p.ptrs, p.R, p.n_levels, p.fmt, C.byref(params), 0, g.H, tile_stride, tile_phase,
                             frame.ctypes.data, C.byref(st))
    assert rc == 0
    return frame, st


def render_frame(scene, view, proj, R: int, W: int, H: int, params: TraceParams | None = None, n_levels: int = 7, fmt: int = FMT_RGBA8):
    base, vst = voxelize(scene, R, accum_mode=ACCUM_FP16 if fmt == FMT_RGBA16F else ACCUM_ORDERED)
    pyr = mipmap(base, n_levels, fmt)
    g = gbuffer(scene, view, proj, W, H)
    frame, tst = trace(scene, view, g, pyr, params)
    return dict(base=base, pyramid=pyr, gbuffer=g, frame=frame, voxel_stats=vst, trace_stats=tst)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    """OpenMP threads of the oracle (torchrun exports OMP_NUM_THREADS=1 to its ranks: bench.py's CPU legs ask for all the cores)"""
    lib().orc_set_num_threads(int(n))


def debug_set_lod_filter(mode: int) -> None:
    """TEST SWITCH: 0 = rule R7, 1 = Mesa llvmpipe's brilinear mip filter (tests/test_gl_llvmpipe.py restores 0)."""
    lib().orc_debug_set_lod_filter(int(mode))


def debug_set_unorm_unpack(mode: int) -> None:
    """TEST SWITCH: 0 = rule R6 (unorm8 -> float = c / 255), 1 = c * (1.0f / 255.0f), as Mesa llvmpipe converts texels (tests/test_gl_llvmpipe.py)."""
    lib().orc_debug_set_unorm_unpack(int(mode))



def debug_set_mip_balanced_sum(on: int) -> None:
    """TEST SWITCH: 1 = the mip filter's four terms added as a balanced tree, as Mesa's GLSL compiler arranges them (tests/test_gl_llvmpipe.py)."""
    lib().orc_debug_set_mip_balanced_sum(int(on))


def render_forward(scene, view, proj, p: Pyramid, W: int, H: int, params: TraceParams | None = None):
    """TEST-ONLY: the visualisation pass as a forward renderer (every fragment that passes the depth test when it is drawn is shaded and
    blended over what is there) -- what a GL pipeline does; gbuffer() + trace() keep the nearest fragment only."""
    sr = SceneRef(scene)
    params = params or default_params()
    frame = np.zeros((H, W), np.uint32)
    rc = lib().orc_render_forward(C.byref(sr.c), _fp(view), _fp(proj), W, H, p.ptrs, p.R, p.n_levels, C.byref(params), frame.ctypes.data)
    assert rc == 0
    return frame
