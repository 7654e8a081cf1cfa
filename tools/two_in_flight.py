"""Experiment: frame throughput with TWO frames in flight on one GPU -- two independent pipelines (device objects = stream sets, grids,
targets) rendering alternate frames, so that the clear -> voxelize -> mip chain of frame i+1 overlaps the cone kernel of frame i.
Public C ABI only; wall clock over 200 frames, frame read-back not included."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxel_cone_tracing_b200 import capi, scene as S

R, W, H = 256, 1920, 1080
sc = S.cornell_scene()
view, proj = S.reference_camera(W / H)
prm = capi.default_params(sampler=1)
pipes = [capi.Pipeline(sc, R, W, H) for _ in range(2)]
for n_pipes in (1, 2):
    use = pipes[:n_pipes]
    for _ in range(4):
        for p in use: p.render_frame(view, proj, prm)
    for p in use: p.sync()
    N = 200
    t0 = time.perf_counter()
    for i in range(N):
        p = use[i % n_pipes]
        p.scene.upload(sc)
        p.render_frame(view, proj, prm)
    for p in use: p.sync()
    dt = time.perf_counter() - t0
    print(f"{n_pipes} frame(s) in flight: {1e3 * dt / N:.3f} ms/frame ({N / dt:.0f} frames/s)", flush=True)
for p in pipes: p.close()
