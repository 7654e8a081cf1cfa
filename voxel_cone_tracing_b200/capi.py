"""ctypes binding of libvct_cuda.so (include/vct/vct_c.h) and a thin object wrapper that mirrors
the reference's Renderer call sequence (src/renderer.h:124-155).  NO CPU fallback: a missing
library or a missing GPU raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import scene as S

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvct_cuda.so")

EXPORTS = [
    "vct_debug_frame_events", "vct_device_create", "vct_device_destroy", "vct_device_sync", "vct_device_stream", "vct_last_error", "vct_version",
    "vct_scene_create", "vct_scene_destroy", "vct_scene_set_geometry", "vct_scene_set_materials", "vct_scene_set_draws",
    "vct_scene_set_lights", "vct_scene_set_cube_size",
    "vct_grid_create", "vct_grid_create_ex", "vct_grid_download_f16", "vct_grid_destroy", "vct_grid_clear", "vct_grid_upload_base", "vct_grid_download",
    "vct_grid_base_device_ptr", "vct_grid_bytes", "vct_grid_occupancy_words", "vct_grid_download_occupancy", "vct_grid_download_array",
    "vct_target_create", "vct_target_destroy", "vct_target_download_frame", "vct_target_download_frame_async", "vct_target_download_wait",
    "vct_target_download_gbuffer", "vct_target_frame_device_ptr",
    "vct_voxelize", "vct_voxelize_reserve", "vct_voxelize_stats", "vct_voxelize_set_accum_mode", "vct_mipmap", "vct_gbuffer", "vct_cone_trace", "vct_cone_trace_count",
    "vct_render_frame", "vct_last_frame_timings", "vct_debug_set",
    "vct_peer_export", "vct_peer_connect", "vct_peer_disconnect", "vct_peer_error",
    "vct_tex3d_create", "vct_tex3d_destroy", "vct_tex3d_clear", "vct_tex3d_mip", "vct_tex3d_upload", "vct_tex3d_download",
]


class VctError(RuntimeError):
    pass


class TraceParams(C.Structure):
    _fields_ = [("enable_direct", C.c_int32), ("enable_diffuse", C.c_int32), ("enable_specular", C.c_int32), ("enable_shadow", C.c_int32),
                ("view_voxel_dir", C.c_int32), ("view_voxel_lod", C.c_float), ("n_diffuse_cones", C.c_int32),
                ("tile_rank", C.c_int32), ("tile_nranks", C.c_int32), ("sampler", C.c_int32)]


class VoxelStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("fragments", "occupied", "items", "capacity", "max_per_voxel")]


class TraceStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("shaded_pixels", "samples_diffuse", "samples_shadow", "samples_specular", "samples_refraction")]

    @property
    def samples(self):
        return self.samples_diffuse + self.samples_shadow + self.samples_specular + self.samples_refraction


def default_params(**kw) -> TraceParams:
    p = TraceParams(1, 1, 1, 1, 7, 0.0, 9, 0, 1, DEFAULT_SAMPLER)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


_lib = None
PEER_HANDLE_BYTES = 320   # sizeof(vct_peer_handle_t)
DEBUG_MIP_DENSE, DEBUG_CONE_VARIANT, DEBUG_CONE_GRID, DEBUG_CONE_RESERVE_SMS, DEBUG_TRACE_LOW_PRIORITY, DEBUG_PEER_REPLICATE, DEBUG_CONE_CTAS_PER_SM, DEBUG_SMALL_LIMIT = 1, 2, 3, 4, 5, 6, 7, 8   # vct_debug_set keys
ACCUM_ORDERED, ACCUM_FIXED_POINT = 0, 1                          # vct_voxelize_set_accum_mode
GRID_RGBA8, GRID_RGBA16F = 0, 1                                  # vct_grid_create_ex formats
SAMPLER_FP32, SAMPLER_TEX = 0, 1
DEFAULT_SAMPLER = int(os.environ.get("VCT_SAMPLER", "0"))


def load():
    """Loads libvct_cuda.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VctError(f"{LIB_PATH} is missing: build it with `make` (nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, u32, f32p = C.c_void_p, C.c_int, C.c_uint32, C.POINTER(C.c_float)
    L.vct_last_error.restype = C.c_char_p
    L.vct_version.restype = C.c_char_p
    L.vct_device_create.argtypes = [i32, C.POINTER(vp)]
    L.vct_device_destroy.argtypes = [vp]
    L.vct_device_sync.argtypes = [vp]
    L.vct_debug_frame_events.argtypes = [vp, vp, C.POINTER(C.c_float)]
    L.vct_device_stream.argtypes = [vp]; L.vct_device_stream.restype = vp
    L.vct_scene_create.argtypes = [vp, C.POINTER(vp)]
    L.vct_scene_destroy.argtypes = [vp]
    L.vct_scene_set_geometry.argtypes = [vp, vp, u32, vp, u32]
    L.vct_scene_set_materials.argtypes = [vp, vp, u32]
    L.vct_scene_set_draws.argtypes = [vp, vp, u32]
    L.vct_scene_set_lights.argtypes = [vp, vp, u32]
    L.vct_scene_set_cube_size.argtypes = [vp, C.c_float]
    L.vct_grid_create.argtypes = [vp, i32, i32, C.POINTER(vp)]
    L.vct_grid_create_ex.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.vct_grid_download_f16.argtypes = [vp, C.c_int, C.c_int, vp]
    L.vct_grid_destroy.argtypes = [vp]
    L.vct_grid_clear.argtypes = [vp]
    L.vct_grid_upload_base.argtypes = [vp, vp]
    L.vct_grid_download.argtypes = [vp, i32, i32, vp]
    L.vct_grid_base_device_ptr.argtypes = [vp]; L.vct_grid_base_device_ptr.restype = vp
    L.vct_grid_bytes.argtypes = [vp]; L.vct_grid_bytes.restype = C.c_size_t
    L.vct_grid_occupancy_words.argtypes = [vp, i32, i32]; L.vct_grid_occupancy_words.restype = C.c_size_t
    L.vct_grid_download_occupancy.argtypes = [vp, i32, i32, vp]
    L.vct_grid_download_array.argtypes = [vp, i32, i32, vp]
    L.vct_target_create.argtypes = [vp, i32, i32, C.POINTER(vp)]
    L.vct_target_destroy.argtypes = [vp]
    L.vct_target_download_frame.argtypes = [vp, vp]
    L.vct_target_download_frame_async.argtypes = [vp, vp, C.POINTER(C.c_uint64)]
    L.vct_target_download_wait.argtypes = [vp, C.c_uint64]
    L.vct_target_download_gbuffer.argtypes = [vp, vp, vp, vp, vp, vp]
    L.vct_target_frame_device_ptr.argtypes = [vp]; L.vct_target_frame_device_ptr.restype = vp
    L.vct_voxelize.argtypes = [vp, vp, vp, i32, i32]
    L.vct_voxelize_reserve.argtypes = [vp, C.c_uint64]
    L.vct_voxelize_stats.argtypes = [vp, C.POINTER(VoxelStats)]
    L.vct_voxelize_set_accum_mode.argtypes = [vp, C.c_int]
    L.vct_mipmap.argtypes = [vp, vp]
    L.vct_gbuffer.argtypes = [vp, vp, f32p, f32p, vp]
    L.vct_cone_trace.argtypes = [vp, vp, vp, f32p, C.POINTER(TraceParams), vp]
    L.vct_cone_trace_count.argtypes = [vp, vp, vp, f32p, C.POINTER(TraceParams), vp, C.POINTER(TraceStats)]
    L.vct_render_frame.argtypes = [vp, vp, vp, vp, f32p, f32p, C.POINTER(TraceParams)]
    L.vct_last_frame_timings.argtypes = [vp, f32p]
    L.vct_debug_set.argtypes = [vp, i32, i32]
    L.vct_peer_export.argtypes = [vp, vp, vp, vp]
    L.vct_peer_connect.argtypes = [vp, vp, vp, i32, i32, vp, i32]
    L.vct_peer_disconnect.argtypes = [vp]
    L.vct_peer_error.argtypes = [vp]
    L.vct_tex3d_create.argtypes = [vp, i32, i32, i32, i32, C.POINTER(vp)]
    L.vct_tex3d_destroy.argtypes = [vp]
    L.vct_tex3d_clear.argtypes = [vp, f32p]
    L.vct_tex3d_mip.argtypes = [vp]
    L.vct_tex3d_upload.argtypes = [vp, i32, vp]
    L.vct_tex3d_download.argtypes = [vp, i32, vp]
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise VctError(f"vct error {rc}: {load().vct_last_error().decode()}")


def _f32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class Device:
    """class Device (src/device.h:32-38) given a body: CUDA device + stream + arenas."""

    def __init__(self, ordinal: int = 0):
        self.L = load()
        self.h = C.c_void_p()
        check(self.L.vct_device_create(ordinal, C.byref(self.h)))

    def sync(self):
        check(self.L.vct_device_sync(self.h))

    def set_accum_mode(self, mode: int):
        """ACCUM_ORDERED (the reference's running average, default) or ACCUM_FIXED_POINT (order-independent integer mean; vct_c.h)"""
        check(self.L.vct_voxelize_set_accum_mode(self.h, mode))

    def debug_set(self, key: int, value: int):
        """measurement / test switches (vct_debug_set): DEBUG_MIP_DENSE, DEBUG_CONE_VARIANT, DEBUG_CONE_GRID"""
        check(self.L.vct_debug_set(self.h, key, value))

    def close(self):
        if self.h:
            self.L.vct_device_destroy(self.h); self.h = C.c_void_p()


class Grid:
    def __init__(self, dev: Device, R: int, levels: int = 7, fmt: int = 0):
        self.dev, self.R, self.levels, self.fmt = dev, R, levels, fmt
        self.h = C.c_void_p()
        check(dev.L.vct_grid_create_ex(dev.h, R, levels, fmt, C.byref(self.h)))

    def clear(self): check(self.dev.L.vct_grid_clear(self.h))

    def upload_base(self, base: np.ndarray):
        b = np.ascontiguousarray(base, np.uint32)
        assert b.size == self.R ** 3
        check(self.dev.L.vct_grid_upload_base(self.h, b.ctypes.data))

    def download(self, level: int, d: int = 0) -> np.ndarray:
        n = self.R >> level
        out = np.empty((n, n, n), np.uint32)
        check(self.dev.L.vct_grid_download(self.h, level, d, out.ctypes.data))
        return out

    def download_f16(self, level: int, d: int = 0) -> np.ndarray:
        """RGBA16F grids: (R >> level)^3 texels as uint64 = four halves (R in the low 16 bits)"""
        n = self.R >> level
        out = np.empty((n, n, n), np.uint64)
        check(self.dev.L.vct_grid_download_f16(self.h, level, d, out.ctypes.data))
        return out

    def download_array(self, level: int, d: int) -> np.ndarray:
        """level >= 1 as stored in the mipmapped CUDA array the texture units read"""
        n = self.R >> level
        out = np.empty((n, n, n), np.uint32)
        check(self.dev.L.vct_grid_download_array(self.h, level, d, out.ctypes.data))
        return out

    def occupancy(self, level: int, dilated: bool) -> np.ndarray:
        """occupancy bits of one level as a bool volume: [N,N,N] (plain) or [N+1,N+1,N+1] (dilated, index = coordinate + 1)"""
        n = self.R >> level
        words = np.empty(int(self.dev.L.vct_grid_occupancy_words(self.h, level, int(dilated))), np.uint32)
        check(self.dev.L.vct_grid_download_occupancy(self.h, level, int(dilated), words.ctypes.data))
        bits = np.unpackbits(words.view(np.uint8), bitorder="little")
        if not dilated:
            return bits[:n ** 3].reshape(n, n, n).astype(bool)
        wpr = (n + 32) // 32
        return bits.reshape(n + 1, n + 1, wpr * 32)[:, :, :n + 1].astype(bool)

    @property
    def base_ptr(self) -> int: return int(self.dev.L.vct_grid_base_device_ptr(self.h))

    @property
    def nbytes(self) -> int: return int(self.dev.L.vct_grid_bytes(self.h))

    def close(self):
        if self.h:
            self.dev.L.vct_grid_destroy(self.h); self.h = C.c_void_p()


class Target:
    def __init__(self, dev: Device, W: int, H: int):
        self.dev, self.W, self.H = dev, W, H
        self.h = C.c_void_p()
        check(dev.L.vct_target_create(dev.h, W, H, C.byref(self.h)))

    def frame(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.H, self.W), np.uint32)
        check(self.dev.L.vct_target_download_frame(self.h, out.ctypes.data))
        return out

    def frame_async(self, out: np.ndarray) -> int:
        """start the read-back of the finished frame into `out` (pinned host memory) on the copy stream; returns a ticket for wait()"""
        assert out.nbytes == self.W * self.H * 4 and out.flags.c_contiguous
        tk = C.c_uint64(0)
        check(self.dev.L.vct_target_download_frame_async(self.h, out.ctypes.data, C.byref(tk)))
        return int(tk.value)

    def wait(self, ticket: int):
        check(self.dev.L.vct_target_download_wait(self.h, ticket))

    def gbuffer(self):
        H, W = self.H, self.W
        tri = np.empty((H, W), np.uint32); depth = np.empty((H, W), np.float32)
        pos = np.empty((H, W, 3), np.float32); nrm = np.empty((H, W, 3), np.float32); mat = np.empty((H, W), np.uint32)
        check(self.dev.L.vct_target_download_gbuffer(self.h, tri.ctypes.data, depth.ctypes.data, pos.ctypes.data, nrm.ctypes.data, mat.ctypes.data))
        return dict(tri_id=tri, depth=depth, world_pos=pos, normal=nrm, material=mat)

    @property
    def frame_ptr(self) -> int: return int(self.dev.L.vct_target_frame_device_ptr(self.h))

    def close(self):
        if self.h:
            self.dev.L.vct_target_destroy(self.h); self.h = C.c_void_p()


class DeviceScene:
    """Device-side copy of a scene.Scene."""

    def __init__(self, dev: Device, sc: S.Scene | None = None):
        self.dev = dev
        self.h = C.c_void_p()
        check(dev.L.vct_scene_create(dev.h, C.byref(self.h)))
        if sc is not None:
            self.upload(sc)

    def upload(self, sc: S.Scene, geometry: bool = True):
        L = self.dev.L
        self.keep = sc
        if geometry:
            v = np.ascontiguousarray(sc.verts); i = np.ascontiguousarray(sc.indices, np.uint32)
            check(L.vct_scene_set_geometry(self.h, v.ctypes.data, len(v), i.ctypes.data, len(i)))
        m = np.ascontiguousarray(sc.materials)
        check(L.vct_scene_set_materials(self.h, m.ctypes.data, len(m)))
        d = np.ascontiguousarray(sc.draws)
        check(L.vct_scene_set_draws(self.h, d.ctypes.data, len(d)))
        l = np.ascontiguousarray(sc.lights)
        check(L.vct_scene_set_lights(self.h, l.ctypes.data if len(l) else None, len(l)))
        check(L.vct_scene_set_cube_size(self.h, float(sc.cube_size)))

    def close(self):
        if self.h:
            self.dev.L.vct_scene_destroy(self.h); self.h = C.c_void_p()


def screen_tile_owner(tx, ty, nranks: int):
    """rank that shades the 32x32 screen tile (tx, ty) of a frame split over `nranks` (screen_tile_owner in csrc/vct_internal.cuh); numpy arrays welcome"""
    k = 3 if nranks % 3 else (5 if nranks % 5 else 7)
    return (tx + k * ty) % nranks


class Pipeline:
    """voxelize -> mip -> G-buffer -> trace, the sequence of Renderer::render() (src/renderer.cpp:392-405)."""

    def __init__(self, sc: S.Scene, R: int, W: int, H: int, levels: int = 7, ordinal: int = 0, reserve: int | None = None, fmt: int = 0):
        self.dev = Device(ordinal)
        self.scene = DeviceScene(self.dev, sc)
        self.grid = Grid(self.dev, R, levels, fmt)
        self.target = Target(self.dev, W, H)
        if reserve:
            check(self.dev.L.vct_voxelize_reserve(self.dev.h, reserve))

    def clear(self): self.grid.clear()

    def voxelize(self, z0: int = 0, z1: int | None = None):
        check(self.dev.L.vct_voxelize(self.dev.h, self.scene.h, self.grid.h, z0, self.grid.R if z1 is None else z1))

    def voxel_stats(self) -> VoxelStats:
        st = VoxelStats()
        check(self.dev.L.vct_voxelize_stats(self.dev.h, C.byref(st)))
        return st

    def mipmap(self): check(self.dev.L.vct_mipmap(self.dev.h, self.grid.h))

    def gbuffer(self, view, proj):
        v = np.ascontiguousarray(view, np.float32); p = np.ascontiguousarray(proj, np.float32)
        check(self.dev.L.vct_gbuffer(self.dev.h, self.scene.h, _f32p(v), _f32p(p), self.target.h))

    def trace(self, view, params: TraceParams | None = None):
        v = np.ascontiguousarray(view, np.float32); params = params or default_params()
        check(self.dev.L.vct_cone_trace(self.dev.h, self.scene.h, self.grid.h, _f32p(v), C.byref(params), self.target.h))

    def trace_count(self, view, params: TraceParams | None = None) -> TraceStats:
        v = np.ascontiguousarray(view, np.float32); params = params or default_params(); st = TraceStats()
        check(self.dev.L.vct_cone_trace_count(self.dev.h, self.scene.h, self.grid.h, _f32p(v), C.byref(params), self.target.h, C.byref(st)))
        return st

    def render_frame(self, view, proj, params: TraceParams | None = None):
        v = np.ascontiguousarray(view, np.float32); p = np.ascontiguousarray(proj, np.float32); params = params or default_params()
        check(self.dev.L.vct_render_frame(self.dev.h, self.scene.h, self.grid.h, self.target.h, _f32p(v), _f32p(p), C.byref(params)))

    def timings(self) -> dict:
        t = np.zeros(8, np.float32)
        check(self.dev.L.vct_last_frame_timings(self.dev.h, _f32p(t)))
        return dict(zip(("clear", "voxelize", "mipmap", "gbuffer", "trace", "total", "cone_kernel", "gbuffer_pass"), [float(x) for x in t[:8]]))

    def sync(self): self.dev.sync()

    # ---- multi-GPU: exchange over NVLink peer memory (include/vct/vct_c.h "multi-GPU") ----
    def peer_export(self) -> bytes:
        """this rank's handle (send it to every other rank by any transport)"""
        buf = C.create_string_buffer(PEER_HANDLE_BYTES)
        check(self.dev.L.vct_peer_export(self.dev.h, self.grid.h, self.target.h, buf))
        return buf.raw

    def peer_connect(self, rank: int, nranks: int, handles: list, frame_root: int = 0):
        """handles[r] = peer_export() of rank r.  Follow with a process barrier before the first render_frame()."""
        assert len(handles) == nranks and all(len(h) == PEER_HANDLE_BYTES for h in handles)
        blob = C.create_string_buffer(b"".join(handles), PEER_HANDLE_BYTES * nranks)
        check(self.dev.L.vct_peer_connect(self.dev.h, self.grid.h, self.target.h, rank, nranks, blob, frame_root))

    def peer_check(self): check(self.dev.L.vct_peer_error(self.dev.h))

    def peer_disconnect(self): check(self.dev.L.vct_peer_disconnect(self.dev.h))

    def close(self):
        self.target.close(); self.grid.close(); self.scene.close(); self.dev.close()
