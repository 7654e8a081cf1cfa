"""Development helper: time the cone kernel of the current VCT_CONE_VARIANT and print a checksum of the frame."""
import sys, os, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from voxel_cone_tracing_b200 import capi, scene as S

cfgs = ((128, 512, 512, False), (256, 1920, 1080, False), (512, 2560, 1440, True))
for (R, W, H, suz) in cfgs:
    sc = S.cornell_scene(with_suzanne=suz)
    view, proj = S.reference_camera(W / H)
    p = capi.Pipeline(sc, R, W, H)
    for sampler in (1, 0):
        nr = int(os.environ.get("TILE_NRANKS", "1"))   # emulate one rank of an N-GPU tile split on one GPU
        prm = capi.default_params(sampler=sampler, tile_rank=0, tile_nranks=nr)
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
        fr = p.target.frame()
        print(f"variant={os.environ.get('VCT_CONE_VARIANT','-')} R={R} {W}x{H} suz={suz} sampler={sampler}: cone={acc['cone_kernel']*1000:.1f}us trace={acc['trace']*1000:.1f}us total={acc['total']*1000:.1f}us crc={zlib.crc32(fr.tobytes()):08x}", flush=True)
    p.close()
