"""Development helper: where the end-to-end loop of bench.py spends its time (wall clock per frame, 200 frames each)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from voxel_cone_tracing_b200 import capi, scene as S

R, W, H = 256, 1920, 1080
sc = S.cornell_scene()
view, proj = S.reference_camera(W / H)
p = capi.Pipeline(sc, R, W, H)
prm = capi.default_params(sampler=1)
host = [torch.empty((H, W), dtype=torch.int32).pin_memory().numpy().view(np.uint32) for _ in range(2)]
N = 200

def run(name, upload, readback):
    for _ in range(3):
        p.render_frame(view, proj, prm)
    p.sync()
    t0 = time.perf_counter(); cpu = 0.0
    prev = None
    for i in range(N):
        c0 = time.perf_counter()
        if upload: p.scene.upload(sc)
        p.render_frame(view, proj, prm)
        if readback == "async":
            tk = p.target.frame_async(host[i & 1])
        cpu += time.perf_counter() - c0
        if readback == "async":
            if prev is not None: p.target.wait(prev)
            prev = tk
        elif readback == "blocking":
            p.target.frame(host[0])
    if prev is not None: p.target.wait(prev)
    p.sync()
    dt = time.perf_counter() - t0
    print(f"{name:40s} {1e3*dt/N:.3f} ms/frame   host enqueue time {1e3*cpu/N:.3f} ms/frame", flush=True)

mode = "-"
run(f"[mode {mode}] render only", False, None)
run(f"[mode {mode}] render + async readback", False, "async")
run(f"[mode {mode}] upload + render + async readback", True, "async")
p.close()
