"""Sweep of the set-up kernels' small-triangle limit (VCT_DEBUG_SMALL_LIMIT) on a synthetic scene: per-stage CUDA-event times.
   python tools/small_limit_sweep.py [config=5]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from voxel_cone_tracing_b200 import capi, scene as S

cfg = bench.CONFIGS[int(sys.argv[1]) if len(sys.argv) > 1 else 5]
sc = bench.build_scene(cfg)
R, W, H = cfg["R"], cfg["W"], cfg["H"]
view, proj = S.reference_camera(W / H)
prm = capi.default_params(sampler=1, n_diffuse_cones=5, enable_diffuse=0, enable_specular=0, enable_shadow=0)   # the trace is not what is measured
p = capi.Pipeline(sc, R, W, H, reserve=24 * sc.n_triangles)
for lim in [int(x) for x in os.environ.get("LIMS", "-1,0,4,9,16,25,36,64,100,144").split(",")]:
    p.dev.debug_set(capi.DEBUG_SMALL_LIMIT, lim)
    acc = {}
    for i in range(5):
        p.render_frame(view, proj, prm)
        if i >= 2:
            for k, v in p.timings().items():
                acc[k] = acc.get(k, 0.0) + v * 1e3 / 3
    print(f"small limit {lim:4d}: voxelize {acc['voxelize']:9.1f} us   G-buffer pass {acc['gbuffer_pass']:9.1f} us", flush=True)
p.close()
