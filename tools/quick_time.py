"""Quick per-stage timing of the benchmark frame (development helper, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from voxel_cone_tracing_b200 import capi, scene as S

for (R, W, H, suz) in ((128, 512, 512, False), (256, 1920, 1080, False), (512, 2560, 1440, True)):
    sc = S.cornell_scene(with_suzanne=suz)
    view, proj = S.reference_camera(W / H)
    p = capi.Pipeline(sc, R, W, H)
    for sampler in (0, 1):
        prm = capi.default_params(sampler=sampler)
        for _ in range(3):
            p.render_frame(view, proj, prm)
        p.sync()
        acc = {}
        n = 10
        for _ in range(n):
            p.render_frame(view, proj, prm)
            t = p.timings()
            for k, v in t.items():
                acc[k] = acc.get(k, 0.0) + v / n
        cnt = p.trace_count(view, prm)
        st = p.voxel_stats()
        print(f"R={R} {W}x{H} suzanne={suz} sampler={sampler}: " + " ".join(f"{k}={v*1000:.1f}us" for k, v in acc.items()),
              f"| samples={cnt.samples/1e6:.1f}M ({cnt.samples/acc['trace']/1e6:.2f} Gsamples/s) frags={st.fragments} items={st.items} occ={st.occupied}", flush=True)
    p.close()
