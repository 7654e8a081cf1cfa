#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native voxel cone tracing hot path.

Metric (BASELINE.json): frames/s of a full frame = clear + revoxelize + six-direction mip build +
G-buffer + cone trace, on configs[1]: CornellBox-Glossy, 256^3 grid, 1920x1080, 9 diffuse + 1 specular
+ 1 shadow cone per pixel, 1 light.  One "step" = one frame.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1..5] [--no-extra]

N > 1 is launched by the driver under torchrun (one rank per GPU): voxelization is sharded by z-slab and
every rank stores its resolved voxels into the peers' grids over NVLink, every rank builds the mip chain
locally, cone tracing is split by 32x32 screen tiles whose pixels are stored into the root's frame.
Besides the headline config the line carries `extra_configs`: BASELINE.json's configs 4 (1 M triangles,
512^3, 3840x2160) and 5 (4 M triangles, 1024^3, 7680x4320, 16 cones) device-timed at the same N, so that
the driver's 1/2/4/8 runs hold the scaling curves north_star asks for.

`--impl reference` times the reference's own GLSL compiled for the CPU (oracle/_ref/libvct_glsl_ref.so, kind "reference"; no
OpenGL 4.5 driver exists in this image -- see DESIGN.md) or, where that library is absent or cannot render the workload, the CPU
oracle (kind "port"), on all host cores, on a bounded sample of the same workload.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from voxel_cone_tracing_b200 import scene as S  # noqa: E402

CONFIGS = {
    1: dict(name="CornellBox-Glossy 128^3 512x512", R=128, W=512, H=512, scene="cornell"),
    2: dict(name="CornellBox-Glossy 256^3 1920x1080", R=256, W=1920, H=1080, scene="cornell"),
    3: dict(name="CornellBox+Suzanne 512^3 2560x1440 (dynamic object)", R=512, W=2560, H=1440, scene="cornell+suzanne"),
    4: dict(name="synthetic 1M triangles 512^3 3840x2160", R=512, W=3840, H=2160, scene="synthetic", tris=1_000_012, seed=0x5EED0001),
    5: dict(name="synthetic 4M triangles 1024^3 7680x4320 (RGBA8, 7 levels; 16 diffuse cones: BASELINE config 5's cone variant)", R=1024, W=7680, H=4320,
            scene="synthetic", tris=4_000_000, seed=0x5EED0002, cones=16),
}
# kernels per frame (one GPU), checked against the ncu launch list (profiles/r02_launches_c2.csv): clear 1 (sparse: the previous frame's occupied
# voxels) + voxelize 4 (counter reset, setup+scan, raster, resolve) + mip 2 (levels 1-3 streaming; tail = levels 4.. + occupancy dilation)
# + G-buffer 4 (clear + counter resets, setup+scan, raster, resolve + live-tile list) + trace 2 (cones, shade).
# N > 1 adds the flag kernels and the tile push: signal, wait for the destination frame, frame push (+ the root's wait for everybody's tiles)
KERNELS_PER_FRAME = 13
KERNELS_PER_FRAME_MULTI = 16
# ncu --set full counters of the dominant kernel on the headline workload, extracted by tools/ncu_to_json.py from a capture of THIS tree
CONE_PROFILE = os.path.join(ROOT, "profiles", "r02_cone_kernel_ncu.json")


def build_scene(cfg, frame: int = 0):
    if cfg["scene"] == "cornell":
        return S.cornell_scene()
    if cfg["scene"] == "cornell+suzanne":
        return S.cornell_scene(with_suzanne=True, theta=0.05 * frame)
    return S.synthetic_scene(cfg["tris"], cfg["seed"])


def config_dict(cfg, world: int, sampler: int, exchange: str, n_tris: int, fif: int = 1) -> dict:
    """identical for both arms (the driver compares them)"""
    R, W, H = cfg["R"], cfg["W"], cfg["H"]
    par = f"z-slab voxelize + screen-tile trace x{world}"
    if world > 1 and exchange == "p2p":
        par = ((f"every rank voxelizes the whole scene (<= 16 k triangles: no voxel exchange) + screen-tile trace x{world}" if n_tris <= 16384 else
                f"z-slab voxelize with sparse voxel push over NVLink peer memory fused into the resolve kernel + screen-tile trace x{world}")
               + ", finished tiles stored into the root's frame over NVLink (CUDA IPC, epoch flags, no collective)")
    elif world > 1:
        par += ", NCCL all-gather of the base level + all-reduce of the frame"
    ws = (7.43 * R ** 3 + W * H * 40) / 1e6
    return {"workload": cfg["name"] + ", revoxelize+mip+gbuffer+trace per frame, %d diffuse + 1 specular + 1 shadow cone" % cfg.get("cones", 9),
            "sampler": "texture units (levels >= 1), software level 0" if sampler == 1 else "software fp32 trilinear",
            "grid": R, "frame": [W, H], "triangles": n_tris, "parallelism": par,
            "frames_in_flight": fif,
            "l2": "no explicit flush: pyramid + G-buffer + frame working set (%.0f MB) exceeds the 126 MB L2 and is rewritten every frame" % ws}


def frames_in_flight(args, n_tris: int, world: int, exchange: str) -> int:
    """--frames-in-flight 0 (default) = automatic: two pipelines for scenes whose front half (clear, voxelize, mip, G-buffer) is latency-bound
    launch chains (<= 64 k triangles: the reference's scenes), one for the large synthetic scenes, where that half is real whole-GPU work
    and a second frame beside the cone kernel only adds contention (config 5 on one GPU: 91.1 ms with one, 95.7 ms with two)"""
    if world > 1 and exchange != "p2p":
        return 1
    return args.frames_in_flight or (2 if n_tris <= 65536 else 1)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU oracle legs
class CpuWorkload:
    """The same frame on the host cores.  Two back ends:
    * kind "reference": THE REFERENCE'S OWN GLSL executed on the CPU (oracle/_ref/libvct_glsl_ref.so -- its six shaders rewritten
      syntactically and compiled against its vendored GLM, build "glm" = GLM's own built-ins; OpenMP over rows / slices), used when
      the library is there and the workload is one the reference can render (RGBA8, 9 diffuse cones);
    * kind "port": the restated oracle (C++/OpenMP) otherwise.
    voxelize + mip + G-buffer are run (and timed) in full once; a "step" then traces every `stride`-th 32x32 screen tile (a
    different phase each step).  With stride 1 a step IS the full trace; with stride > 1 the frame time is t_voxelize + t_mip +
    t_gbuffer + stride * t_trace_step and the line says that it is extrapolated."""

    def __init__(self, cfg):
        from oracle import orc
        self.orc, self.cfg = orc, cfg
        self.ref = None
        if cfg.get("cones", 9) == 9:
            try:
                from oracle import glsl_ref
                if glsl_ref.available():
                    glsl_ref.lib()
                    self.ref = glsl_ref
            except (OSError, subprocess.CalledProcessError):
                self.ref = None
        self.kind = "reference" if self.ref else "port"
        # torchrun exports OMP_NUM_THREADS=1 to every rank: the baseline uses all the cores whoever launched it
        orc.set_num_threads(os.cpu_count() or 1)
        self.sc = build_scene(cfg)
        R, W, H = cfg["R"], cfg["W"], cfg["H"]
        self.view, self.proj = S.reference_camera(W / H)
        orc.mipmap(np.zeros((8, 8, 8), np.uint32), 4)   # spin up the OpenMP pool outside the timed part
        if self.ref:
            t0 = time.perf_counter(); tex, _ = self.ref.voxelize(self.sc, R, "glm")
            t1 = time.perf_counter(); self.pyr = self.ref.mipmap(tex[0], 7, "glm")
            t2 = time.perf_counter(); self.g = self.ref.gbuffer(self.sc, self.view, self.proj, W, H, "glm")
        else:
            t0 = time.perf_counter(); base, _ = orc.voxelize(self.sc, R)
            t1 = time.perf_counter(); self.pyr = orc.mipmap(base, 7)
            t2 = time.perf_counter(); self.g = orc.gbuffer(self.sc, self.view, self.proj, W, H)
        t3 = time.perf_counter()
        self.t_vox, self.t_mip, self.t_gbuf = t1 - t0, t2 - t1, t3 - t2
        self.cores = orc.num_threads()
        self.frame = np.zeros((H, W), np.uint32)
        self.n_tiles = ((W + 31) // 32) * ((H + 31) // 32)

    def trace_step(self, stride: int, phase: int):
        """-> (seconds, samples taken or None: the reference's shader does not count them)"""
        t0 = time.perf_counter()
        if self.ref:
            self.ref.shade(self.sc, self.view, self.g, self.pyr, None, stride, phase % stride, "glm")
            return time.perf_counter() - t0, None
        _, st = self.orc.trace(self.sc, self.view, self.g, self.pyr, self.orc.default_params(n_diffuse_cones=self.cfg.get("cones", 9)), stride, phase % stride, self.frame)
        dt = time.perf_counter() - t0
        return dt, int(st.samples)

    def pick_stride(self, steps: int, budget_s: float) -> int:
        """smallest power-of-two thinning so that `steps` steps fit in the budget"""
        dt, _ = self.trace_step(64, 0)              # calibration: 1/64 of the tiles
        full = dt * 64.0
        per_step = max(budget_s / max(steps, 1), 0.05)
        stride = 1
        while full / stride > per_step and stride < self.n_tiles // 4:
            stride *= 2
        return stride

    def frame_seconds(self, t_trace_step: float, stride: int) -> float:
        return self.t_vox + self.t_mip + self.t_gbuf + t_trace_step * stride

    def sample_text(self, stride: int) -> str:
        how = ("each step cone-traces the whole frame" if stride == 1 else
               f"each step cone-traces every {stride}th 32x32 tile; the frame time is EXTRAPOLATED: voxelize + mip + G-buffer + {stride} x the step's trace time")
        what = ("the reference's own GLSL (shader/*.vert|geom|frag|comp rewritten syntactically, compiled against its vendored GLM, oracle/_ref/libvct_glsl_ref.so) "
                "executed on the CPU behind the oracle's fixed-function rules" if self.ref else "oracle = CPU restatement of the GLSL")
        return (f"{what} (C++/OpenMP, {self.cores} threads; no OpenGL 4.5 driver in this image: the reference executable cannot run): "
                f"voxelize+mip+G-buffer of '{self.cfg['name']}' in full (timed once), {how}")


def run_reference(args, cfg, rank: int, world: int):
    if rank != 0:
        return
    wl = CpuWorkload(cfg)
    stride = wl.pick_stride(args.steps + args.warmup, 100.0)
    for i in range(args.warmup):
        wl.trace_step(stride, i)
    ts = [wl.trace_step(stride, args.warmup + i)[0] for i in range(args.steps)]
    frame_ms = 1e3 * sum(wl.frame_seconds(t, stride) for t in ts) / len(ts)
    # what one step really took on the wall: the shared stages are run once, outside the steps
    step_ms = frame_ms if stride == 1 else 1e3 * sum(ts) / len(ts)
    v = 1e3 / frame_ms
    print(json.dumps({
        "impl": "reference", "metric": "frames_per_sec", "value": v, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "frame_ms": frame_ms, "extrapolated": stride != 1, "tile_stride": stride,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 (u8 RGBA storage)", "data": "synthetic",
        "config": config_dict(cfg, world, args.sampler, args.exchange, wl.sc.n_triangles, frames_in_flight(args, wl.sc.n_triangles, world, args.exchange)),
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": wl.cores, "kind": wl.kind, "sample": wl.sample_text(stride),
                         "voxelize_s": wl.t_vox, "mip_s": wl.t_mip, "gbuffer_s": wl.t_gbuf},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


# ----------------------------------------------------------------------------- GPU arm
class CudaArray:
    """__cuda_array_interface__ view of device memory owned by libvct_cuda (for torch.distributed)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 3, "strides": None}


class Rig:
    """one config on this rank's GPU: pipeline + the frame function of the chosen exchange"""

    def __init__(self, cfg, args, rank, world, local_rank, torch, dist, fp16: bool = False):
        from voxel_cone_tracing_b200 import capi
        self.capi, self.torch, self.dist = capi, torch, dist
        self.cfg, self.args, self.rank, self.world, self.local_rank = cfg, args, rank, world, local_rank
        R, W, H = cfg["R"], cfg["W"], cfg["H"]
        self.sc = build_scene(cfg)
        self.view, self.proj = S.reference_camera(W / H)
        levels = (R.bit_length() if fp16 else 7)   # fp16 variant: the full chain, log2(R) + 1 levels
        self.prm = capi.default_params(tile_rank=rank, tile_nranks=world, sampler=args.sampler, n_diffuse_cones=cfg.get("cones", 9))
        self.z0, self.z1 = rank * R // world, (rank + 1) * R // world
        self.p2p = world > 1 and args.exchange == "p2p"
        # Frames in flight: F independent pipelines (device object = stream set + arenas, grid, target) render alternate frames.  The
        # front half of frame i+1 (clear, voxelize, exchange, mip, G-buffer: chains of small latency-bound kernels) then runs beside the
        # trace of frame i; the traces of all pipelines are serialised on one low-priority stream per GPU (VCT_DEBUG_TRACE_LOW_PRIORITY).
        self.F = frames_in_flight(args, self.sc.n_triangles, world, args.exchange)
        self.pipes, self.streams = [], []
        for k in range(self.F):
            pipe = capi.Pipeline(self.sc, R, W, H, levels, ordinal=local_rank, reserve=max(1 << 20, 8 * self.sc.n_triangles),
                                 fmt=capi.GRID_RGBA16F if fp16 else capi.GRID_RGBA8)
            if self.F > 1 and args.trace_stream == "shared":
                pipe.dev.debug_set(capi.DEBUG_TRACE_LOW_PRIORITY, 1)
            if args.replicate >= 0:
                pipe.dev.debug_set(capi.DEBUG_PEER_REPLICATE, args.replicate)
            if args.cone_ctas:
                pipe.dev.debug_set(capi.DEBUG_CONE_CTAS_PER_SM, args.cone_ctas)
            if args.cone_grid:
                pipe.dev.debug_set(capi.DEBUG_CONE_GRID, 1)
            if args.reserve_sms:
                pipe.dev.debug_set(capi.DEBUG_CONE_RESERVE_SMS, args.reserve_sms)
            self.pipes.append(pipe)
            self.streams.append(torch.cuda.ExternalStream(int(pipe.dev.L.vct_device_stream(pipe.dev.h)), device=torch.device("cuda", local_rank)))
            # size the fragment arena for this rank's slab before anything is timed (the library grows it on overflow and asks for a re-run)
            for _ in range(4):
                pipe.clear(); pipe.voxelize(self.z0, self.z1)
                try:
                    pipe.voxel_stats()
                    break
                except capi.VctError as e:
                    if "overflow" not in str(e):
                        raise
            pipe.clear()
        self.pipe, self.stream = self.pipes[0], self.streams[0]
        self.n_frames = 0
        pipe = self.pipe
        self.base_t = self.frame_t = None
        if world > 1 and not self.p2p:
            self.base_t = torch.as_tensor(CudaArray(pipe.grid.base_ptr, (R * R * R,), "<i4"), device=torch.device("cuda", local_rank))
            self.frame_t = torch.as_tensor(CudaArray(pipe.target.frame_ptr, (W * H,), "<i4"), device=torch.device("cuda", local_rank))
        if self.p2p:
            # NVLink peer-memory exchange fused into the resolve / shade kernels (csrc/peer.cu): handles travel once, here (one connection per pipeline)
            for pipe in self.pipes:
                handles = [None] * world
                dist.all_gather_object(handles, pipe.peer_export())
                pipe.peer_connect(rank, world, handles, frame_root=0)
            dist.barrier()

    def next_pipe(self):
        pipe = self.pipes[self.n_frames % self.F]
        self.n_frames += 1
        return pipe

    def frame(self):
        pipe, dist, torch = self.pipe, self.dist, self.torch
        if self.world == 1 or self.p2p:
            self.next_pipe().render_frame(self.view, self.proj, self.prm)
            return
        R = self.cfg["R"]
        per_rank = R * R * R // self.world
        with torch.cuda.stream(self.stream):
            pipe.clear()
            pipe.voxelize(self.z0, self.z1)
            dist.all_gather_into_tensor(self.base_t, self.base_t[self.rank * per_rank:(self.rank + 1) * per_rank])   # in place, z-major slabs
            pipe.mipmap()
            pipe.gbuffer(self.view, self.proj)
            self.frame_t.zero_()
            pipe.trace(self.view, self.prm)
            dist.all_reduce(self.frame_t)   # tiles are disjoint: integer sum == merge

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        for pipe in self.pipes:
            pipe.sync()
        self.torch.cuda.synchronize()

    def timed(self, steps: int, warmup: int) -> float:
        """ms per frame, device-timed with CUDA events on the launching stream, max over ranks"""
        torch = self.torch
        for _ in range(max(warmup, 3)):
            self.frame()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record(self.stream)
        for _ in range(steps):
            self.frame()
        for st in self.streams[1:]:   # the clock stops when the last frame of EVERY pipeline is done
            ev = torch.cuda.Event()
            ev.record(st)
            self.stream.wait_event(ev)
        e1.record(self.stream)
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], device=f"cuda:{self.local_rank}")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    def stage_times(self, n: int) -> dict:
        """per-stage device times of this rank in us (CUDA events inside vct_render_frame), averaged over n frames"""
        acc = {}
        for _ in range(n):   # one pipeline, one frame at a time: the stages of a frame on an otherwise idle GPU
            self.pipe.render_frame(self.view, self.proj, self.prm)
            for k, v in self.pipe.timings().items():
                acc[k] = acc.get(k, 0.0) + v * 1e3 / n
        return acc

    def close(self):
        if self.p2p:
            for pipe in self.pipes:
                pipe.peer_check()         # raises if a flag wait ever timed out
            self.barrier()                # nobody unmaps while a peer may still be storing into it
            for pipe in self.pipes:
                pipe.peer_disconnect()
            self.barrier()
        for pipe in self.pipes:
            pipe.close()


def h2d_bytes(sc) -> int:
    return int(sc.verts.nbytes + sc.indices.nbytes + sc.materials.nbytes + sc.draws.nbytes + sc.lights.nbytes + 128 + 36)


def e2e_loop(rig: Rig, steps: int):
    """the frame through the public C ABI with HOST buffers: every step uploads the scene (geometry, materials, draw list, lights)
    from host memory and reads the finished frame back into pinned host memory; wall clock, max over ranks"""
    torch, sc, rank, world = rig.torch, rig.sc, rig.rank, rig.world
    W, H = rig.cfg["W"], rig.cfg["H"]
    F = rig.F
    # ring of 3F pinned frames: the copy of frame i out of its pipeline's snapshot is started by that pipeline's NEXT frame (i + F, see
    # vct_target_download_frame_async), so the host collects frame i - 2F while frame i is being enqueued and never waits on work it has only just queued
    NB = 3 * F
    host_frames = [torch.empty((H, W), dtype=torch.int32).pin_memory().numpy().view(np.uint32) for _ in range(NB)] if rank == 0 else None
    for i in range(2 * F):
        pipe = rig.next_pipe()
        pipe.scene.upload(sc); pipe.render_frame(rig.view, rig.proj, rig.prm)
        if rank == 0:
            pipe.target.wait(pipe.target.frame_async(host_frames[i]))
    rig.barrier()
    t0 = time.perf_counter()
    pending = []                                             # (pipeline, ticket) of the read-backs in flight, oldest first
    for i in range(steps):
        pipe = rig.next_pipe()
        pipe.scene.upload(sc)                                # H2D: geometry, materials, draw list (pinned staging ring, async)
        pipe.render_frame(rig.view, rig.proj, rig.prm)
        if rank == 0:
            if len(pending) >= 2 * F:
                q, tk = pending.pop(0)
                q.target.wait(tk)                             # frame i-2F is in host memory (its buffer is reused by frame i+F)
            pending.append((pipe, pipe.target.frame_async(host_frames[i % NB])))   # D2H of this frame, asynchronous: overlaps the next frames
    for q, tk in pending:
        q.target.wait(tk)
    for pipe in rig.pipes:
        pipe.sync()
    dt = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([dt], device=f"cuda:{rig.local_rank}")
        rig.dist.all_reduce(tt, op=rig.dist.ReduceOp.MAX)
        dt = float(tt.item())
    out = {"value": steps / dt, "unit": "frames/s", "h2d_bytes_per_step": h2d_bytes(sc) * world, "d2h_bytes_per_step": int(W * H * 4),
           "ms_per_step": 1e3 * dt / steps,
           "note": ("every rank uploads the scene from host memory and renders its share every step, the root reads the merged frame back to pinned host "
                    "memory" if world > 1 else "scene uploaded from host memory and the finished frame read back to pinned host memory EVERY step through the C ABI")
                   + "; the read-back of frame i overlaps the rendering of frame i+1 (vct_target_download_frame_async); wall clock"}
    if world == 1:   # the same loop fully serialised (one pipeline, blocking read-back each step), for reference
        n = min(steps, 50)
        pipe = rig.pipe
        pipe.sync()
        t0 = time.perf_counter()
        for i in range(n):
            pipe.scene.upload(sc); pipe.render_frame(rig.view, rig.proj, rig.prm); pipe.target.frame(host_frames[0])
        pipe.sync()
        out["blocking_ms_per_step"] = 1e3 * (time.perf_counter() - t0) / n
    return out


def cone_roofline(rig: Rig, cone_us: float, sm_mhz, cnt):
    """The dominant kernel against the unit that binds it.  The pyramid it gathers from is L1/L2 resident (DRAM traffic is < 1 % of the
    HBM peak), so an HBM roofline says nothing; the binding unit is the texture pipe: one TEX wavefront per clock per SM.  The number
    of wavefronts of one launch is a property of the workload (same frame, same kernel) and is read from the ncu capture of this tree
    (path and hash printed); the time is the live CUDA-event time of this run."""
    peak_hbm, peak_src = measured_peaks()
    out = {"kernel": "cone_kernel_fast", "bound": "l1tex", "unit": "Gwavefronts/s", "achieved": None, "peak": None, "frac": None, "traffic": None,
           "samples_per_launch": int(cnt.samples), "gsamples_per_s": cnt.samples / (cone_us * 1e-6) / 1e9, "kernel_us": cone_us,
           "hbm_peak_gbs": peak_hbm, "hbm_peak_source": peak_src}
    key = f"config{rig.args.config}_sampler{rig.args.sampler}"
    try:
        raw = open(CONE_PROFILE, "rb").read()
        prof = json.loads(raw)[key]
    except Exception as ex:
        out["note"] = f"no ncu capture for {key} in {os.path.relpath(CONE_PROFILE, ROOT)} ({type(ex).__name__}): only Gsamples/s is reported"
        return out
    wf = float(prof["tex_wavefronts"])
    sms = 148
    clk = (sm_mhz or 1965.0) * 1e6
    dram = float(prof["dram_bytes_read"]) + float(prof["dram_bytes_write"])
    out.update({
        "achieved": wf / (cone_us * 1e-6) / 1e9, "peak": sms * clk / 1e9, "frac": wf / (cone_us * 1e-6) / (sms * clk),
        "traffic": dram, "tex_wavefronts_per_launch": wf, "warp_instructions_per_launch": prof.get("warp_instructions"),
        "hbm_frac": dram / (cone_us * 1e-6) / 1e9 / peak_hbm,
        "ncu": {"file": os.path.relpath(CONE_PROFILE, ROOT), "sha256": hashlib.sha256(raw).hexdigest()[:16], "kernel_us_under_ncu": prof.get("time_us"),
                "tex_pipe_pct_under_ncu": prof.get("tex_wavefront_pct")},
        "note": "achieved = TEX wavefronts of one launch (ncu capture of this tree, same workload) / CUDA-event kernel time of this run; "
                "peak = 148 SMs x 1 wavefront per clock x the SM clock sampled during the run; traffic = DRAM bytes per launch (ncu)"})
    return out


def mip_stage(rig: Rig, stage_acc: dict) -> dict:
    """mip stage against its HBM roofline (7.4286 R^3 algorithmic bytes).  The roofline fraction is the DENSE build
    (every tile read and written); the running frame loop moves far less (untouched all-zero tiles are skipped) and is reported in us."""
    torch, pipe, capi = rig.torch, rig.pipe, rig.capi
    R = rig.cfg["R"]
    peak, peak_src = measured_peaks()
    alg = 7.4286 * R ** 3

    def dense_us(n=10):
        pipe.dev.debug_set(capi.DEBUG_MIP_DENSE, 1)
        pipe.mipmap(); pipe.sync()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record(rig.stream)
        for _ in range(n):
            pipe.mipmap()
        d1.record(rig.stream)
        pipe.sync()
        pipe.dev.debug_set(capi.DEBUG_MIP_DENSE, 0)
        return d0.elapsed_time(d1) * 1e3 / n

    scene_us = dense_us()
    out = {"bound": "hbm", "unit": "GB/s", "peak": peak, "peak_source": peak_src, "algorithmic_bytes": alg,
           "achieved": alg / (scene_us * 1e-6) / 1e9, "frac": alg / (scene_us * 1e-6) / 1e9 / peak, "dense_us": scene_us,
           "frame_loop_us": stage_acc["mipmap"],
           "note": "frac = dense algorithmic bytes / time of the DENSE build of this frame's grid (every tile read, every output written; "
                   "vct_debug_set(VCT_DEBUG_MIP_DENSE)); frame_loop_us = the stage inside the running frame loop, where untouched all-zero tiles are skipped"}
    if R <= 256:   # the same build on a uniform-random grid (SURVEY 8(d) micro-input): every texel takes the arithmetic path
        g2 = capi.Grid(pipe.dev, R, 7)
        rng = np.random.default_rng(1)
        g2.upload_base(rng.integers(0, 2 ** 32, (R, R, R), dtype=np.uint64).astype(np.uint32))
        keep = pipe.grid
        pipe.grid = g2
        rnd = dense_us()
        pipe.grid = keep
        g2.close()
        out["dense_random_us"] = rnd
        out["dense_random_frac"] = alg / (rnd * 1e-6) / 1e9 / peak
    pipe.render_frame(rig.view, rig.proj, rig.prm); pipe.sync()      # back to the tracked state
    return out


def run_extra(cfg_id, args, rank, world, local_rank, torch, dist, fp16: bool = False):
    """one of the large configs, device-timed at this N (no end-to-end leg, no CPU leg)"""
    cfg = CONFIGS[cfg_id]
    steps = 10 if cfg_id == 4 else (3 if fp16 else 5)
    try:
        rig = Rig(cfg, args, rank, world, local_rank, torch, dist, fp16)
        ms = rig.timed(steps, 3)
        st = rig.stage_times(3) if (world == 1 or rig.p2p) else None
        per_rank = None
        if world > 1 and st is not None:
            per_rank = [None] * world
            dist.all_gather_object(per_rank, {k: round(v, 1) for k, v in st.items()})
        wl = config_dict(cfg, world, args.sampler, args.exchange, rig.sc.n_triangles, rig.F)["workload"]
        if fp16:
            wl = wl.replace("RGBA8, 7 levels", "RGBA16F grid, full chain of %d levels: BASELINE config 5's storage variant, fp32 software filtering" % cfg["R"].bit_length())
        out = {"workload": wl, "ms_per_frame": ms, "frames_per_s": 1e3 / ms,
               "steps": steps, "grid": cfg["R"], "frame": [cfg["W"], cfg["H"]], "triangles": rig.sc.n_triangles}
        if world == 1:
            out["stages_us"] = {k: round(v, 1) for k, v in st.items()}
            out["fragments"] = int(rig.pipe.voxel_stats().fragments)
        elif per_rank:
            out["stages_us_per_rank"] = per_rank
        rig.close()
        return out
    except Exception as ex:   # never lose the headline line to an extra config
        return {"error": f"{type(ex).__name__}: {ex}"}


def run_ours(args, cfg, rank: int, world: int, local_rank: int):
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rig = Rig(cfg, args, rank, world, local_rank, torch, dist, fp16=args.fp16)
    pipe = rig.pipe
    # ---- timed region: exactly K frames after >= 3 warm-up frames, CUDA events on the launching stream, max over ranks ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step = rig.timed(args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    value = 1e3 / ms_step
    st = pipe.voxel_stats()

    stage_acc, rank_stages = None, None
    if world == 1 or rig.p2p:
        stage_acc = rig.stage_times(min(args.steps, 20))
        if world > 1:
            rank_stages = [None] * world
            dist.all_gather_object(rank_stages, {k: round(v, 1) for k, v in stage_acc.items()})

    e2e = None
    if world == 1 or rig.p2p:
        try:
            e2e = e2e_loop(rig, args.steps)
        except Exception as ex:   # never lose the device-timed line to the end-to-end leg
            if rank == 0:
                print(f"e2e leg failed: {ex}", file=sys.stderr)

    roof, stages = None, None
    if world == 1 and not args.fp16:
        cnt = pipe.trace_count(rig.view, rig.prm)
        roof = cone_roofline(rig, stage_acc["cone_kernel"], (clocks or {}).get("sm_mhz"), cnt)
        stages = {k + "_us": v for k, v in stage_acc.items()}
        stages["mip_roofline"] = mip_stage(rig, stage_acc)
        stages["note"] = ("gbuffer_us = the part of the G-buffer pass on the critical path: the pass (gbuffer_pass_us) runs on a second stream "
                          "beside clear + voxelize + mip and joins before the trace, so the stage times overlap and need not add up to total_us; "
                          "clear_us = the sparse clear of the previous frame's occupied voxels (occupied_voxels words), not a bandwidth figure")
        stages["voxelize_mfrag_per_s"] = st.fragments / (stage_acc["voxelize"] * 1e-6) / 1e6
        stages["fragments"] = int(st.fragments)
        stages["occupied_voxels"] = int(st.occupied)
        stages["shaded_pixels"] = int(cnt.shaded_pixels)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and not args.fp16:
        wl = CpuWorkload(cfg)
        stride = wl.pick_stride(4, 16.0)
        ts = [wl.trace_step(stride, i)[0] for i in range(4)]
        fs = sum(wl.frame_seconds(t, stride) for t in ts) / len(ts)
        cpu = {"value": 1.0 / fs, "unit": "frames/s", "cores": wl.cores, "kind": wl.kind, "sample": wl.sample_text(stride) + " (4 steps)",
               "frame_s": fs, "voxelize_s": wl.t_vox, "mip_s": wl.t_mip, "gbuffer_s": wl.t_gbuf}

    n_tris, n_fif = rig.sc.n_triangles, rig.F
    rig.close()
    extra = None
    if not args.no_extra and args.config == 2 and not args.fp16 and (world == 1 or args.exchange == "p2p"):
        extra = {str(c): run_extra(c, args, rank, world, local_rank, torch, dist) for c in (4, 5)}
        if world == 1:   # config 5 as BASELINE.json states it: fp16 RGBA grid + full mip chain (single GPU: the exchange is RGBA8 only)
            extra["5_fp16_full_chain"] = run_extra(5, args, rank, world, local_rank, torch, dist, fp16=True)

    if rank == 0:
        out = {"metric": "frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
               "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 (u8 RGBA storage)",
               "data": "synthetic", "config": dict(config_dict(cfg, world, args.sampler, args.exchange, n_tris, n_fif), **({"storage": "RGBA16F, full mip chain (variant)"} if args.fp16 else {})),
               "clocks": clocks, "e2e": e2e, "gpu_launches": (KERNELS_PER_FRAME if world == 1 else KERNELS_PER_FRAME_MULTI) * args.steps, "roofline": roof, "cpu_baseline": cpu}
        if stages:
            out["stages"] = stages
        if rank_stages:
            out["stages_us_per_rank"] = rank_stages
        if extra:
            out["extra_configs"] = extra
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--fp16", action="store_true", help="run the chosen config with the RGBA16F grid + full mip chain storage variant (one GPU; not the headline)")
    ap.add_argument("--frames-in-flight", type=int, default=0, choices=[0, 1, 2, 3],
                    help="independent pipelines rendering alternate frames (the front half of frame i+1 runs beside the trace of frame i); 1 = the plain loop, 0 = automatic (2 for scenes of <= 64 k triangles)")
    ap.add_argument("--trace-stream", default="shared", choices=["shared", "own"],
                    help="frames in flight: cones + shade of all pipelines on one low-priority stream per GPU (default) or on each pipeline's own stream")
    ap.add_argument("--replicate", type=int, default=-1, choices=[-1, 0, 1],
                    help="multi-GPU voxelization: 1 = every rank voxelizes the whole scene, 0 = z-slabs + voxel push, -1 = by scene size (library default)")
    ap.add_argument("--cone-ctas", type=int, default=0, help="experiment: CTAs per SM of the persistent cone kernel (default: all that fit, 10)")
    ap.add_argument("--cone-grid", action="store_true", help="experiment: cone kernel on a host-sized grid instead of the persistent work queue")
    ap.add_argument("--reserve-sms", type=int, default=0, help="experiment: SMs the persistent cone kernel leaves to the other pipeline's front half")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the device-timed runs of configs 4 and 5 (extra_configs)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU exchange: p2p = sparse voxel push + tile push over NVLink peer memory, fused into the kernels (default); "
                         "nccl = dense in-place all-gather of the base level + all-reduce of the frame (the library baseline)")
    ap.add_argument("--sampler", type=int, default=1, choices=[0, 1],
                    help="textureLod evaluator of the cone tracer: 1 = texture units (default; frame within 2/255, PSNR > 60 dB of the oracle), 0 = software fp32")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return
    run_ours(args, cfg, rank, world, local_rank)


if __name__ == "__main__":
    main()
