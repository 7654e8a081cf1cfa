"""Experiment: frame throughput with TWO frames in flight on one GPU -- two independent pipelines (device objects = stream sets, grids,
targets) rendering alternate frames, so that the clear -> voxelize -> mip -> G-buffer chain of frame i+1 overlaps the cone kernel of
frame i.  Variants: persistent cone kernel / host-sized grid.  Results and the stream-priority prototype: profiles/r02_frames_in_flight.txt.
Public C ABI only; wall clock over N frames, scene upload every frame, no read-back."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxel_cone_tracing_b200 import capi, scene as S

R, W, H = 256, 1920, 1080
sc = S.cornell_scene()
view, proj = S.reference_camera(W / H)
prm = capi.default_params(sampler=1)
pipes = [capi.Pipeline(sc, R, W, H) for _ in range(3)]
cases = [(1, 0, 0, 0), (2, 0, 0, 0), (1, 0, 0, 1)] + [(2, 0, k, 1) for k in (0, 2, 4, 6, 8, 10, 12, 16, 24)] + [(3, 0, 8, 1), (3, 0, 4, 1)]
for n_pipes, grid, reserve, low in cases:
    use = pipes[:n_pipes]
    for p in use:
        p.dev.debug_set(capi.DEBUG_CONE_GRID, grid)
        p.dev.debug_set(capi.DEBUG_CONE_RESERVE_SMS, reserve)
        p.dev.debug_set(capi.DEBUG_TRACE_LOW_PRIORITY, low)
    for _ in range(4):
        for p in use: p.render_frame(view, proj, prm)
    for p in use: p.sync()
    N = 300
    t0 = time.perf_counter()
    for i in range(N):
        p = use[i % n_pipes]
        p.scene.upload(sc)
        p.render_frame(view, proj, prm)
    for p in use: p.sync()
    dt = time.perf_counter() - t0
    print(f"{n_pipes} frame(s) in flight, cone_grid={grid}, reserved SMs={reserve}, trace on low-priority stream={low}: {1e3 * dt / N:.3f} ms/frame ({N / dt:.0f} frames/s)", flush=True)
    print("   last frame of pipeline 0 (us):", {k: round(v * 1e3, 1) for k, v in use[0].timings().items()}, flush=True)
for p in pipes: p.close()
