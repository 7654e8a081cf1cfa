#!/usr/bin/env python
"""Writes tests/golden/glsl_ref_vectors.json: outputs of THE REFERENCE'S OWN GLSL executed on the CPU (oracle/glsl_ref/, mode
"rules": built-ins evaluated by the written rules R5 / R9, fixed-function stages by R1-R4, R6-R8), on seeded inputs.

The library needs /root/reference (shader text + GLM, read where they lie) and cannot be rebuilt on a box without it; these vectors
travel instead, so tests/test_glsl_ref_golden.py can hold the restated oracle against reference-sourced results anywhere.

    python tools/make_glsl_ref_golden.py        (needs /root/reference; runs `make -C oracle ref`)
"""
import base64
import hashlib
import json
import os
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from oracle import glsl_ref as G  # noqa: E402
from voxel_cone_tracing_b200 import scene as S  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "glsl_ref_vectors.json")


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def pack(a) -> str:
    return base64.b64encode(zlib.compress(np.ascontiguousarray(a).tobytes(), 9)).decode()


def main():
    rng = np.random.default_rng(2026)
    out = {"source": "oracle/_ref/libvct_glsl_ref.so mode rules = /root/reference/shader/*.{vert,geom,frag,comp} @ fb8d2717 over thirdparty/glm 0.9.9",
           "generator": "tools/make_glsl_ref_golden.py"}

    # imageAtomicRGBA8Avg, voxelize.frag:95-120: sequences of 20 values folded into one texel
    folds = []
    for s in range(8):
        vals = rng.random((20, 4)).astype(np.float32) if s % 2 else (np.round(rng.random((20, 4)) * 8) / 8).astype(np.float32)
        stored, res = 0, []
        for v in vals:
            stored = G.fold(stored, v)
            res.append(stored)
        folds.append({"vals_f32_hex": vals.view(np.uint32).reshape(-1).tolist(), "stored": res})
    out["fold_sequences"] = folds

    # axis selection, voxelize.geom:25-55
    tris = rng.standard_normal((48, 3, 3)).astype(np.float32)
    tris[40] = [[0.11, 0.23, 0.37], [1.11, -0.77, 0.87], [0.36, -0.02, 1.37]]      # |n.x| == |n.y| ties
    tris[41] = [[0.11, 0.23, 0.37], [1.11, 0.73, -0.63], [0.36, 1.23, 0.12]]
    out["axis"] = {"tris_f32_hex": tris.view(np.uint32).reshape(-1).tolist(),
                   "axis": [G.select_axis(t[0], t[1], t[2]) for t in tris]}

    # voxel grids + mip chains
    grids = []
    for (R, suz, theta) in [(32, True, 0.4), (64, True, 0.3), (128, False, 0.0)]:
        sc = S.cornell_scene(with_suzanne=suz, theta=theta)
        tex, n = G.voxelize(sc, R)
        assert all(np.array_equal(tex[i], tex[0]) for i in range(6))
        levels = 7 if R >= 64 else 6
        pyr = G.mipmap(tex[0], levels)
        g = {"R": R, "suzanne": suz, "theta": theta, "levels": levels, "fragments_executed": n, "occupied": int((tex[0] != 0).sum()),
             "base_sha256": sha(tex[0]),
             "mip_sha256": {f"{d}.{l}": sha(pyr.levels[d][l]) for d in range(6) for l in range(1, levels)}}
        if R == 32:   # the small grid in full: index / value pairs, so that a failure can say which voxel
            idx = np.flatnonzero(tex[0])
            g["voxels"] = {"index": idx.tolist(), "value": tex[0].reshape(-1)[idx].tolist()}
        grids.append(g)
    out["grids"] = grids

    # cones through the (64, True, 0.3) pyramid: trace_cone of voxel_cone_tracing.frag:88-119
    sc = S.cornell_scene(with_suzanne=True, theta=0.3)
    tex, _ = G.voxelize(sc, 64)
    pyr = G.mipmap(tex[0], 7)
    cones = []
    for i in range(64):
        o = rng.random(3).astype(np.float32); d = rng.standard_normal(3).astype(np.float32)
        ap = np.float32([0.55785173935, 0.1, 0.0174533, 1.2][i % 4]); md = np.float32([1.73205080757, 0.7][i % 2])
        r = G.trace_cone(pyr, o, d, float(ap), float(md))
        cones.append({"origin": o.view(np.uint32).tolist(), "dir": d.view(np.uint32).tolist(), "aperture": float(ap), "max_dist": float(md),
                      "rgba_f32_hex": r.view(np.uint32).tolist()})
    out["cones"] = cones

    # G-buffer + frame (Renderer::visualize), camera of main.cpp and one inside the box
    frames = []
    W, H = 160, 120
    for cam in (dict(), dict(eye=(0.2, 0.9, 0.6), pitch=-10.0, yaw=-100.0)):
        view, proj = S.reference_camera(W / H, **cam)
        g = G.gbuffer(sc, view, proj, W, H)
        f = G.shade(sc, view, g, pyr)
        frames.append({"camera": cam, "W": W, "H": H, "tri_id_sha256": sha(g.tri_id), "depth_sha256": sha(g.depth),
                       "world_pos_sha256": sha(np.where((g.tri_id != 0xFFFFFFFF)[..., None], g.world_pos, 0).astype(np.float32)),
                       "normal_sha256": sha(np.where((g.tri_id != 0xFFFFFFFF)[..., None], g.normal, 0).astype(np.float32)),
                       "frame_sha256": sha(f), "frame_zlib_b64": pack(f)})
    out["frames"] = frames

    with open(OUT, "w") as fh:
        json.dump(out, fh, separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
