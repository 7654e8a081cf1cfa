#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native voxel cone tracing hot path.

Metric (BASELINE.json): frames/s of a full frame = clear + revoxelize + six-direction mip build +
G-buffer + cone trace, on configs[1]: CornellBox-Glossy, 256^3 grid, 1920x1080, 9 diffuse + 1 specular
+ 1 shadow cone per pixel, 1 light.  One "step" = one frame.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1..5]

N > 1 is launched by the driver under torchrun (one rank per GPU): voxelization is sharded by z-slab,
the base level is all-gathered in place over NCCL, every rank builds the mip chain locally, cone
tracing is split by 32x32 screen tiles and the frame is merged with one NCCL reduction.

`--impl reference` times the CPU oracle (the reference's GLSL needs an OpenGL 4.5 driver that does not
exist in this image -- see DESIGN.md) on the host cores on a bounded sample of the same workload.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
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
KERNELS_PER_FRAME = 15  # clear 1 (sparse: the previous frame's occupied voxels) + voxelize 4 (counter reset, setup+scan, raster, resolve) + mip 2 (fused low; tail = levels 4-6 + occupancy + dilation) + gbuffer 4 (clear, setup+scan, raster, resolve) + trace 4 (list reset, tile list, cones, shade)
# ncu --set full capture of cone_kernel_fast on this workload (profiles/r01_ncu_s7.md), per launch:
# dram__bytes_read.sum + dram__bytes_write.sum, and the two units that bind the kernel
CONE_KERNEL_DRAM_TRAFFIC = {1: 24.6e6 + 10.4e6}
CONE_KERNEL_NCU = {"tex_wavefront_frac": 0.764, "issue_frac": 0.662, "tex_wavefronts_per_launch": 192.4e6, "warp_instructions_per_launch": 643.4e6,
                   "l1tex_sectors_per_launch": 649.9e6, "source": "profiles/r01_ncu_s7.md (ncu --set full, one launch of this command)"}


def build_scene(cfg, frame: int = 0):
    if cfg["scene"] == "cornell":
        return S.cornell_scene()
    if cfg["scene"] == "cornell+suzanne":
        return S.cornell_scene(with_suzanne=True, theta=0.05 * frame)
    return S.synthetic_scene(cfg["tris"], cfg["seed"])


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
    """The same frame on the host cores with the oracle.  voxelize + mip + G-buffer are run (and timed) in
    full once; a "step" then traces every `stride`-th 32x32 screen tile (a different phase each step) and
    the frame time is estimated as t_voxelize + t_mip + t_gbuffer + stride * t_trace_step."""

    def __init__(self, cfg):
        from oracle import orc
        self.orc, self.cfg = orc, cfg
        self.sc = build_scene(cfg)
        R, W, H = cfg["R"], cfg["W"], cfg["H"]
        self.view, self.proj = S.reference_camera(W / H)
        orc.mipmap(np.zeros((8, 8, 8), np.uint32), 4)   # spin up the OpenMP pool outside the timed part
        t0 = time.perf_counter(); base, _ = orc.voxelize(self.sc, R)
        t1 = time.perf_counter(); self.pyr = orc.mipmap(base, 7)
        t2 = time.perf_counter(); self.g = orc.gbuffer(self.sc, self.view, self.proj, W, H)
        t3 = time.perf_counter()
        self.t_vox, self.t_mip, self.t_gbuf = t1 - t0, t2 - t1, t3 - t2
        self.cores = orc.num_threads()
        self.frame = np.zeros((H, W), np.uint32)
        self.n_tiles = ((W + 31) // 32) * ((H + 31) // 32)

    def trace_step(self, stride: int, phase: int):
        t0 = time.perf_counter()
        _, st = self.orc.trace(self.sc, self.view, self.g, self.pyr, self.orc.default_params(n_diffuse_cones=self.cfg.get("cones", 9)), stride, phase % stride, self.frame)
        dt = time.perf_counter() - t0
        return dt, int(st.samples)

    def pick_stride(self, steps: int, budget_s: float) -> int:
        """largest power-of-two-ish thinning so that `steps` steps fit in the budget"""
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
        return (f"oracle = CPU restatement of the GLSL (C++/OpenMP, {self.cores} threads; Mesa llvmpipe is unavailable in this image): "
                f"voxelize+mip+G-buffer of '{self.cfg['name']}' in full (timed once), each step cone-traces every {stride}th 32x32 tile "
                f"and scales the trace time x{stride}")


def run_reference(args, cfg, rank: int):
    if rank != 0:
        return
    wl = CpuWorkload(cfg)
    stride = wl.pick_stride(args.steps + args.warmup, 100.0)
    for i in range(args.warmup):
        wl.trace_step(stride, i)
    ts = [wl.trace_step(stride, args.warmup + i)[0] for i in range(args.steps)]
    ms = 1e3 * sum(wl.frame_seconds(t, stride) for t in ts) / len(ts)
    v = 1e3 / ms
    print(json.dumps({
        "impl": "reference", "metric": "frames_per_sec", "value": v, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 (u8 RGBA storage)", "data": "synthetic",
        "config": {"workload": cfg["name"] + ", revoxelize+mip+gbuffer+trace per frame, %d diffuse + 1 specular + 1 shadow cone" % cfg.get("cones", 9)},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": wl.cores, "kind": "port", "sample": wl.sample_text(stride),
                         "voxelize_s": wl.t_vox, "mip_s": wl.t_mip, "gbuffer_s": wl.t_gbuf},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


# ----------------------------------------------------------------------------- GPU arm
class CudaArray:
    """__cuda_array_interface__ view of device memory owned by libvct_cuda (for torch.distributed)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 3, "strides": None}


def run_ours(args, cfg, rank: int, world: int, local_rank: int):
    import torch
    from voxel_cone_tracing_b200 import capi
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    R, W, H = cfg["R"], cfg["W"], cfg["H"]
    sc = build_scene(cfg)
    view, proj = S.reference_camera(W / H)
    pipe = capi.Pipeline(sc, R, W, H, 7, ordinal=local_rank, reserve=max(1 << 20, 8 * sc.n_triangles))
    L, dev = pipe.dev.L, pipe.dev
    stream = torch.cuda.ExternalStream(int(L.vct_device_stream(dev.h)), device=torch.device("cuda", local_rank))
    prm = capi.default_params(tile_rank=rank, tile_nranks=world, sampler=args.sampler, n_diffuse_cones=cfg.get("cones", 9))
    z0, z1 = rank * R // world, (rank + 1) * R // world
    base_t = frame_t = None
    if world > 1:
        base_t = torch.as_tensor(CudaArray(pipe.grid.base_ptr, (R * R * R,), "<i4"), device=torch.device("cuda", local_rank))
        frame_t = torch.as_tensor(CudaArray(pipe.target.frame_ptr, (W * H,), "<i4"), device=torch.device("cuda", local_rank))
    per_rank = R * R * R // world
    # size the fragment arena for this rank's slab before anything is timed (the library grows it on overflow and asks for a re-run)
    for _ in range(4):
        pipe.clear(); pipe.voxelize(z0, z1)
        try:
            pipe.voxel_stats()
            break
        except capi.VctError as e:
            if "overflow" not in str(e):
                raise
    pipe.clear()
    p2p = world > 1 and args.exchange == "p2p"
    if p2p:
        # NVLink peer-memory exchange fused into the resolve / shade kernels (csrc/peer.cu): handles travel once, here
        handles = [None] * world
        dist.all_gather_object(handles, pipe.peer_export())
        pipe.peer_connect(rank, world, handles, frame_root=0)
        dist.barrier()

    def frame_device():
        if world == 1 or p2p:
            pipe.render_frame(view, proj, prm)
            return
        with torch.cuda.stream(stream):
            pipe.clear()
            pipe.voxelize(z0, z1)
            dist.all_gather_into_tensor(base_t, base_t[rank * per_rank:(rank + 1) * per_rank])   # in place, z-major slabs
            pipe.mipmap()
            pipe.gbuffer(view, proj)
            frame_t.zero_()
            pipe.trace(view, prm)
            dist.all_reduce(frame_t)   # tiles are disjoint: integer sum == merge

    def barrier():
        if world > 1:
            dist.barrier()
        pipe.sync()
        torch.cuda.synchronize()

    # ---- warm-up, sanity ----
    for _ in range(max(args.warmup, 3)):
        frame_device()
    barrier()
    st = pipe.voxel_stats()
    # ---- timed region: exactly K frames, CUDA events on the launching stream ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_acc = {}
    barrier()
    with torch.cuda.stream(stream):
        e0.record(stream)
    for _ in range(args.steps):
        frame_device()
    with torch.cuda.stream(stream):
        e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    # per-stage device times (events around each stage; separate untimed pass so that the headline has no extra events... they are cheap, but keep it clean)
    n_stage = 0
    if world == 1:
        for _ in range(min(args.steps, 20)):
            pipe.render_frame(view, proj, prm)
            for k, v in pipe.timings().items():
                stage_acc[k] = stage_acc.get(k, 0.0) + v
            n_stage += 1
        stage_acc = {k: v / n_stage for k, v in stage_acc.items()}
    rank_stages = None
    if world > 1:
        t = torch.tensor([ms_total], device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        if p2p:   # per-rank stage times of the last frames (CUDA events inside vct_render_frame), for the scaling analysis
            acc = {}
            for _ in range(10):
                pipe.render_frame(view, proj, prm)
                for k, v in pipe.timings().items():
                    acc[k] = acc.get(k, 0.0) + v * 100.0     # -> us, averaged over 10
            rank_stages = [None] * world
            dist.all_gather_object(rank_stages, {k: round(v, 1) for k, v in acc.items()})
    ms_step = ms_total / args.steps
    value = 1e3 / ms_step

    # ---- end-to-end through the public C-ABI with HOST buffers: per step upload the whole scene from host
    # memory (geometry, materials, draw list, lights) and read the finished frame back into pinned host memory ----
    e2e = None
    if world == 1:
        # two pinned host frames: the read-back of frame i (copy stream) overlaps the rendering of frame i+1 (device stream);
        # every frame's pixels have landed in host memory before the clock stops
        host_frames = [torch.empty((H, W), dtype=torch.int32).pin_memory().numpy().view(np.uint32) for _ in range(2)]
        h2d = sc.verts.nbytes + sc.indices.nbytes + sc.materials.nbytes + sc.draws.nbytes + sc.lights.nbytes + 128 + 36
        d2h = host_frames[0].nbytes
        for i in range(2):
            pipe.scene.upload(sc); pipe.render_frame(view, proj, prm); pipe.target.wait(pipe.target.frame_async(host_frames[i]))
        pipe.sync()
        t0 = time.perf_counter()
        prev = None
        for i in range(args.steps):
            pipe.scene.upload(sc)                                # H2D: geometry, materials, draw list (pinned staging ring, async)
            pipe.render_frame(view, proj, prm)
            tk = pipe.target.frame_async(host_frames[i & 1])     # D2H of this frame, asynchronous
            if prev is not None:
                pipe.target.wait(prev)                           # frame i-1 is in host memory
            prev = tk
        pipe.target.wait(prev)
        pipe.sync()
        dt = time.perf_counter() - t0
        e2e = {"value": args.steps / dt, "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": 1e3 * dt / args.steps,
               "note": "scene uploaded from host memory and the finished frame read back to pinned host memory EVERY step through the C ABI; "
                       "the read-back of frame i overlaps the rendering of frame i+1 (vct_target_download_frame_async)"}
        # the same loop fully serialised (blocking read-back each step), for reference
        pipe.sync()
        t0 = time.perf_counter()
        for i in range(min(args.steps, 50)):
            pipe.scene.upload(sc); pipe.render_frame(view, proj, prm); pipe.target.frame(host_frames[0])
        pipe.sync()
        e2e["blocking_ms_per_step"] = 1e3 * (time.perf_counter() - t0) / min(args.steps, 50)

    if world > 1 and p2p:
        # N GPUs, same loop: every rank uploads the scene and renders its share through the C ABI every step; the root (which receives
        # the other ranks' tiles over NVLink) reads the merged frame back to pinned host memory.  Wall clock, max over ranks.
        try:
            h2d = sc.verts.nbytes + sc.indices.nbytes + sc.materials.nbytes + sc.draws.nbytes + sc.lights.nbytes + 128 + 36
            host_frames = [torch.empty((H, W), dtype=torch.int32).pin_memory().numpy().view(np.uint32) for _ in range(2)] if rank == 0 else None
            for i in range(2):
                pipe.scene.upload(sc); pipe.render_frame(view, proj, prm)
                if rank == 0:
                    pipe.target.wait(pipe.target.frame_async(host_frames[i]))
            barrier()
            t0 = time.perf_counter()
            prev = None
            for i in range(args.steps):
                pipe.scene.upload(sc)
                pipe.render_frame(view, proj, prm)
                if rank == 0:
                    tk = pipe.target.frame_async(host_frames[i & 1])
                    if prev is not None:
                        pipe.target.wait(prev)
                    prev = tk
            if rank == 0 and prev is not None:
                pipe.target.wait(prev)
            pipe.sync()
            tt = torch.tensor([time.perf_counter() - t0], device=f"cuda:{local_rank}")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
            e2e = {"value": args.steps / dt, "unit": "frames/s", "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(W * H * 4),
                   "ms_per_step": 1e3 * dt / args.steps,
                   "note": "every rank uploads the scene from host memory and renders its share every step, the root reads the merged frame back to "
                           "pinned host memory (asynchronously, overlapped with the next frame); wall clock, max over ranks"}
        except Exception as ex:   # never lose the device-timed line to the end-to-end leg
            e2e = None
            if rank == 0:
                print(f"e2e leg failed: {ex}", file=sys.stderr)

    # ---- roofline of the dominant kernel (cone_kernel, timed alone with CUDA events on its stream) ----
    peak, peak_src = measured_peaks()
    roof = None
    stages = None
    if world == 1:
        cnt = pipe.trace_count(view, prm)
        t_trace = stage_acc["cone_kernel"] * 1e-3
        gather_bytes = 192.0 * cnt.samples     # 3 directions x 2 levels x 8 texels x 4 B per sample_voxel
        ach = gather_bytes / t_trace / 1e9
        roof = {"kernel": "cone_kernel", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": CONE_KERNEL_DRAM_TRAFFIC.get(args.sampler) if args.config == 2 else None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": gather_bytes, "samples_per_launch": int(cnt.samples),
                "gsamples_per_s": cnt.samples / t_trace / 1e9,
                "binding_units": CONE_KERNEL_NCU if (args.config == 2 and args.sampler == 1) else None,
                "note": "algorithmic gather bytes (192 B per sample_voxel: 3 directions x 2 levels x 8 texels x 4 B) / CUDA-event kernel time. The gathers are "
                        "served by the texture units / L1 (20.8 GB of L1TEX sectors per launch, 99 % hit) and the 126 MB L2, DRAM traffic is 0.035 GB per launch, "
                        "so frac > 1 against the HBM copy peak is expected; the kernel is bound by the TEX pipe (76 % of one wavefront/clk/SM) and the "
                        "issue slots (66 %), see binding_units and DESIGN.md 3.3"}
        mip_bytes = 7.4286 * R ** 3
        stages = {k + "_us": v * 1e3 for k, v in stage_acc.items()}
        stages["mip_roofline"] = {"bound": "hbm", "achieved": mip_bytes / (stage_acc["mipmap"] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                  "frac": mip_bytes / (stage_acc["mipmap"] * 1e-3) / 1e9 / peak, "algorithmic_bytes": mip_bytes,
                                  "note": "dense algorithmic bytes (7.4286 R^3) / mip stage time of the running frame loop, where tiles that the voxelizer "
                                          "did not touch and whose outputs are already zero are neither read nor written (DESIGN.md 3.2); dense build: "
                                          "mip_dense_us"}
        # the dense mip build (every tile read and written: first frame, uploads), CUDA events on the library's stream
        os.environ["VCT_MIP_DENSE"] = "1"
        pipe.mipmap(); pipe.sync()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            d0.record(stream)
            for _ in range(10):
                pipe.mipmap()
            d1.record(stream)
        pipe.sync()
        del os.environ["VCT_MIP_DENSE"]
        pipe.render_frame(view, proj, prm); pipe.sync()      # back to the tracked state
        stages["mip_dense_us"] = d0.elapsed_time(d1) * 100.0  # ms per 10 builds -> us per build
        stages["mip_roofline"]["dense_frac"] = mip_bytes / (stages["mip_dense_us"] * 1e-6) / 1e9 / peak
        stages["note"] = ("gbuffer_us = the part of the G-buffer pass on the critical path: the pass (gbuffer_pass_us) runs on a second stream "
                          "beside clear + voxelize + mip and joins before the trace, so the stage times overlap and need not add up to total_us")
        stages["clear_gbs"] = 4.0 * R ** 3 / (stage_acc["clear"] * 1e-3) / 1e9
        stages["voxelize_mfrag_per_s"] = st.fragments / (stage_acc["voxelize"] * 1e-3) / 1e6
        stages["fragments"] = int(st.fragments)
        stages["shaded_pixels"] = int(cnt.shaded_pixels)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        wl = CpuWorkload(cfg)
        stride = wl.pick_stride(4, 16.0)
        ts = [wl.trace_step(stride, i)[0] for i in range(4)]
        fs = sum(wl.frame_seconds(t, stride) for t in ts) / len(ts)
        cpu = {"value": 1.0 / fs, "unit": "frames/s", "cores": wl.cores, "kind": "port", "sample": wl.sample_text(stride) + " (4 steps)",
               "frame_s": fs, "voxelize_s": wl.t_vox, "mip_s": wl.t_mip, "gbuffer_s": wl.t_gbuf}

    if rank == 0:
        out = {"metric": "frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
               "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 (u8 RGBA storage)",
               "data": "synthetic",
               "config": {"workload": cfg["name"] + ", revoxelize+mip+gbuffer+trace per frame, %d diffuse + 1 specular + 1 shadow cone" % cfg.get("cones", 9),
                          "sampler": "texture units (levels >= 1), software level 0" if args.sampler == 1 else "software fp32 trilinear",
                          "grid": R, "frame": [W, H], "triangles": sc.n_triangles, "parallelism": f"z-slab voxelize + screen-tile trace x{world}" + ("" if world == 1 else (", sparse NVLink peer-store exchange fused into the resolve/shade kernels (CUDA IPC, no collective)" if p2p else ", NCCL all-gather of the base level + all-reduce of the frame")),
                          "l2": "no explicit flush: grid + G-buffer + frame working set (%.0f MB) exceeds the 126 MB L2 and is rewritten every frame"
                                % ((pipe.grid.nbytes + W * H * 40) / 1e6)},
               "clocks": clocks, "e2e": e2e, "gpu_launches": KERNELS_PER_FRAME * args.steps, "roofline": roof, "cpu_baseline": cpu}
        if stages:
            out["stages"] = stages
        if rank_stages:
            out["stages_us_per_rank"] = rank_stages
        print(json.dumps(out), flush=True)
    if p2p:
        pipe.peer_check()        # raises if a flag wait ever timed out
        barrier()                # nobody unmaps while a peer may still be storing into it
        pipe.peer_disconnect()
        barrier()
    pipe.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU exchange: p2p = sparse voxel push + tile push over NVLink peer memory, fused into the kernels (default); "
                         "nccl = dense in-place all-gather of the base level + all-reduce of the frame (the library baseline)")
    ap.add_argument("--sampler", type=int, default=1, choices=[0, 1],
                    help="textureLod evaluator of the cone tracer: 1 = texture units (default; frame within 2/255, PSNR > 60 dB of the oracle), 0 = software fp32")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return
    run_ours(args, cfg, rank, world, local_rank)


if __name__ == "__main__":
    main()
