#!/bin/bash
# compute-sanitizer memcheck over a small frame sequence (all kernels of the frame path, sparse bookkeeping, async read-back)
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from voxel_cone_tracing_b200 import capi, scene as S
R, W, H = 64, 320, 200
view, proj = S.reference_camera(W / H)
p = capi.Pipeline(S.cornell_scene(with_suzanne=True), R, W, H)
host = [torch.empty((H, W), dtype=torch.int32).pin_memory().numpy().view(np.uint32) for _ in range(2)]
prev = None
for i in range(4):
    p.scene.upload(S.cornell_scene(with_suzanne=True, theta=0.4 * i))
    p.render_frame(view, proj, capi.default_params(sampler=i & 1))
    tk = p.target.frame_async(host[i & 1])
    if prev: p.target.wait(prev)
    prev = tk
p.target.wait(prev)
p.trace_count(view, capi.default_params())
for v in ("0", "1", "2", "3"):
    os.environ["VCT_CONE_VARIANT"] = v
    p.render_frame(view, proj, capi.default_params(sampler=1, n_diffuse_cones=16))
p.grid.upload_base(np.random.default_rng(0).integers(0, 2**32, (R, R, R), dtype=np.uint64).astype(np.uint32)); p.mipmap()
p.render_frame(view, proj); p.sync()
print("frames ok", int(p.target.frame().sum() % 1000003))
p.close()
# a larger frame so that the grouped-diffuse kernel runs
p = capi.Pipeline(S.cornell_scene(), 64, 1920, 1080)
p.render_frame(*S.reference_camera(1920 / 1080), capi.default_params(sampler=1)); p.sync(); p.close()
print("done")
PY
for tool in memcheck racecheck synccheck; do timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python /tmp/san.py > gpurun_out/sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; tail -3 gpurun_out/sanitizer_$tool.log; done
