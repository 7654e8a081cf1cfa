"""How long does the HOST take to enqueue one frame (vct_render_frame through ctypes)?  If this is close to the device time per frame the
loop is launch bound (multi-GPU: 1/8 of the frame per rank).  Prints host microseconds per call and the device time."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxel_cone_tracing_b200 import capi, scene as S

R, W, H = 64, 128, 64   # a frame the device finishes faster than the host can enqueue it: no queue back-pressure in the host time
sc = S.cornell_scene()
view, proj = S.reference_camera(W / H)
for nranks in (1,):
    prm = capi.default_params(sampler=1, tile_rank=0, tile_nranks=nranks)
    p = capi.Pipeline(sc, R, W, H)
    for _ in range(5):
        p.render_frame(view, proj, prm)
    p.sync()
    N = 200
    t0 = time.perf_counter()
    for _ in range(N):
        p.render_frame(view, proj, prm)
    t1 = time.perf_counter()
    p.sync()
    t2 = time.perf_counter()
    print(f"tile share 1/{nranks}: host enqueue {1e6 * (t1 - t0) / N:.1f} us per frame, device {1e6 * (t2 - t0) / N:.1f} us per frame", flush=True)
    t0 = time.perf_counter()
    for _ in range(N):
        p.scene.upload(sc)
    t1 = time.perf_counter()
    p.sync()
    print(f"   scene upload: host {1e6 * (t1 - t0) / N:.1f} us per call", flush=True)
    p.close()
