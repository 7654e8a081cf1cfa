"""Accuracy + speed of the two textureLod evaluators (fp32 software vs texture units) against the oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import orc
from voxel_cone_tracing_b200 import capi, scene as S

def psnr(a, b):
    a = a.view(np.uint8).astype(np.float64); b = b.view(np.uint8).astype(np.float64)
    mse = np.mean((a - b) ** 2)
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)

for (R, W, H, suz) in ((128, 512, 512, True), (256, 1920, 1080, False)):
    sc = S.cornell_scene(with_suzanne=suz)
    view, proj = S.reference_camera(W / H)
    t = time.time(); ref = orc.render_frame(sc, view, proj, R, W, H); print(f"oracle {time.time()-t:.1f}s", flush=True)
    p = capi.Pipeline(sc, R, W, H)
    for sampler in (0, 1):
        prm = capi.default_params(sampler=sampler)
        for _ in range(3):
            p.render_frame(view, proj, prm)
        acc = 0.0
        for _ in range(5):
            p.render_frame(view, proj, prm); acc += p.timings()["trace"] / 5
        got = p.target.frame()
        d = np.abs(got.view(np.uint8).astype(np.int32) - ref["frame"].view(np.uint8).astype(np.int32))
        hist = np.bincount(d.reshape(-1), minlength=6)[:6]
        print(f"R={R} {W}x{H} sampler={sampler}: trace {acc*1000:.0f} us  max_abs={d.max()} psnr={psnr(got, ref['frame']):.2f} dB  hist(0..5)={hist.tolist()}  mip={p.timings()['mipmap']*1000:.0f}us", flush=True)
    p.close()
