#!/usr/bin/env python
"""Model of the texture-unit work of cone_kernel_fast on a workload (not a parity check, not a measurement): the cones of every 2x2
pixel block (= one TEX quad of the kernel) marched on the CPU by the reference's own trace_cone (oracle/_ref/libvct_glsl_ref.so,
run_tex_model in oracle/glsl_ref/harness.cpp), every sample classified the way the kernel classifies it (no fetch / one level / two
levels through the texture unit), and the quad-level fetch units added up for the kernel as it is and for re-mappings of the work.
One unit = one level of one direction for one quad = one TEX wavefront (ncu: 2 per one-level, 4 per two-level request of 8 lanes).

    python tools/tex_lane_model.py [--config 2] [--stride 8]
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
from oracle import glsl_ref as G  # noqa: E402
from oracle import orc  # noqa: E402
from voxel_cone_tracing_b200 import scene as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--stride", type=int, default=8)
    args = ap.parse_args()
    cfg = bench.CONFIGS[args.config]
    sc = bench.build_scene(cfg)
    R, W, H = cfg["R"], cfg["W"], cfg["H"]
    view, proj = S.reference_camera(W / H)
    base, _ = orc.voxelize(sc, R)
    pyr = orc.mipmap(base, 7)
    g = orc.gbuffer(sc, view, proj, W, H)
    L = G.lib()
    fn = L.glref_rules_tex_model
    fn.argtypes = [C.POINTER(orc.SceneT), orc.f32p, C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.c_void_p, C.c_int, C.c_int, C.POINTER(orc.TraceParams),
                                                                                          C.c_int, C.c_int, C.POINTER(C.c_double)]
    sr = orc.SceneRef(sc)
    prm = orc.default_params()
    tot = np.zeros(8)
    out = (C.c_double * 8)()
    rc = fn(C.byref(sr.c), orc._fp(view), W, H, g.tri_id.ctypes.data, g.world_pos.ctypes.data, g.normal.ctypes.data, g.material.ctypes.data,
            pyr.ptrs, R, 7, C.byref(prm), args.stride, 0, out)
    assert rc == 0
    tot = np.array(list(out)) * args.stride
    names = ["ideal (perfectly packed quads)", "kernel as it is (one- and two-level instructions issued separately inside a quad)",
             "quad-uniform level decision", "lane-autonomous march + quad-uniform decision"]
    res = {"workload": cfg["name"], "tile_stride": args.stride, "samples": tot[4], "samples_that_fetch": tot[5],
           "tex_wavefront_units": {n: tot[i] for i, n in enumerate(names)}}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
