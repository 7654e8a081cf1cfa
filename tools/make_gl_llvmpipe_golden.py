#!/usr/bin/env python
"""tests/golden/gl_llvmpipe_frames.npz: frames the REFERENCE's own visualisation shaders (shader/voxel_cone_tracing.vert|frag, read from the
reference tree) rendered on a real OpenGL implementation -- Mesa 18.1 llvmpipe as shipped inside Nsight Compute, driven by
oracle/_ref/gl/vct_gl_ref (oracle/gl_ref/) -- with the voxel textures filled from the oracle's grid + mip chain.  The cases are those of
tests/test_gl_llvmpipe.py (CASES there); the test compares the oracle with these frames on any box and, where llvmpipe and the reference
tree exist, re-renders them.

    python tools/make_gl_llvmpipe_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import gl_ref  # noqa: E402
import test_gl_llvmpipe as T  # noqa: E402


def main():
    if not gl_ref.available():
        sys.exit("needs oracle/_ref/gl/vct_gl_ref (make -C oracle gl), Nsight Compute's Mesa libGL and /root/reference/shader")
    out = {}
    for name in T.CASES:
        sc, view, proj, R, W, H, prm = T.case_inputs(name)
        pyr = T.pyramid(name)
        u8, f32 = gl_ref.visualize(sc, view, proj, pyr, W, H, prm)
        out[name] = u8
        out[name + ":f32"] = f32[::2, ::2].astype(np.float32) if name in T.FLOAT_CASES else np.zeros(0, np.float32)   # every second pixel
        print(name, u8.shape)
    sc, view, proj, R, W, H, prm = T.case_inputs("suzanne")
    out["pipeline_suzanne"] = gl_ref.render_frame(sc, view, proj, R, W, H, prm)["frame"]     # all three passes on the driver
    for kind in ("stack", "outside", "degenerate", "lights", "tir", "mirror", "nolight", "empty"):     # corner cases, all three passes on the driver
        sc, view, proj = T.edge_inputs(kind)
        out["edge:" + kind] = gl_ref.render_frame(sc, view, proj, T.EDGE_R, T.EDGE_W, T.EDGE_H, levels=6)["frame"]
    path = os.path.join(ROOT, "tests", "golden", "gl_llvmpipe_frames.npz")
    np.savez_compressed(path, **out)
    print("->", path, os.path.getsize(path), "bytes")
    vox = {}
    for name in T.VOXEL_CASES:
        sc, res = T.voxel_scene(name)
        tri, xy, voxel, col = gl_ref.voxelize_fragments(sc, res)
        assert np.array_equal(voxel, np.floor(voxel)) and voxel.min() >= 0 and voxel.max() < res
        vox[name + ":tri"], vox[name + ":xy"], vox[name + ":voxel"], vox[name + ":colour"] = tri, xy.astype(np.uint16), voxel.astype(np.int16), col
        print(name, len(tri), "fragments")
    tri, xy, voxel, col = gl_ref.voxelize_fragments(T.raster_soup(), 32)      # per-triangle raster check
    vox["raster_soup:tri"], vox["raster_soup:voxel"] = tri, voxel.astype(np.int16)
    print("raster_soup", len(tri), "fragments")
    np.savez_compressed(T.VOXEL_GOLDEN, **vox)
    print("->", T.VOXEL_GOLDEN, os.path.getsize(T.VOXEL_GOLDEN), "bytes")
    mip = {}
    for name in T.MIP_CASES:
        base = T.mip_base(name)
        levels = int(np.log2(base.shape[0])) + 1
        chain = gl_ref.mip_chain(base, levels)
        for d in range(6):
            for l in range(1, levels):
                mip[f"{name}:{d}:{l}"] = chain[d][l]
        print(name, levels, "levels")
    np.savez_compressed(T.MIP_GOLDEN, **mip)
    print("->", T.MIP_GOLDEN, os.path.getsize(T.MIP_GOLDEN), "bytes")


if __name__ == "__main__":
    main()
