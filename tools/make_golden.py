#!/usr/bin/env python
"""Regenerates tests/golden/*.json.

* tinyobj_streams.json -- SHA-256 of the flattened per-index vertex stream (pos, normal, uv bit patterns
  + material id) and of the material constants, as produced by the REFERENCE's vendored tinyobjloader
  (oracle/_ref/tinyobj_dump, built from /root/reference where it lies).  The CPU tests recompute the same
  stream from the committed assets/*.vctmesh fixtures, so the benchmark inputs stay pinned to what the
  reference's loader reads even on boxes without /root/reference.
* oracle_regression.json -- SHA-256 of oracle outputs on the reference scene (regression guard for the
  oracle itself; NOT a reference pin -- the reference ships no golden data).

    python tools/make_golden.py        (needs /root/reference and `make -C oracle`)
"""
import hashlib
import json
import os
import struct
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from voxel_cone_tracing_b200 import scene as S  # noqa: E402

REF_ASSETS = "/root/reference/assets"
GOLD = os.path.join(ROOT, "tests", "golden")


def tinyobj_stream(obj_path: str):
    """(sha256 of vertex stream, sha256 of materials, counts) from the reference loader's dump."""
    out = subprocess.check_output([os.path.join(ROOT, "oracle", "_ref", "tinyobj_dump"), obj_path], text=True)
    hv, hm = hashlib.sha256(), hashlib.sha256()
    n_idx = n_mat = 0
    for line in out.splitlines():
        tok = line.split()
        if tok[0] == "V":
            hv.update(struct.pack("<8Ii", *[int(t, 16) for t in tok[1:9]], int(tok[9]))); n_idx += 1
        elif tok[0] == "M":
            hm.update(struct.pack("<18Ii", *[int(t, 16) for t in tok[2:20]], int(tok[20]))); n_mat += 1
    return hv.hexdigest(), hm.hexdigest(), n_idx, n_mat


def fixture_stream(mesh: S.Mesh):
    """The same two digests computed from one of our Mesh objects."""
    hv, hm = hashlib.sha256(), hashlib.sha256()
    mat_of_index = np.full(len(mesh.indices), -1, np.int32)
    for (a, b, m) in mesh.ranges:
        mat_of_index[a:a + b] = m
    v = mesh.verts[mesh.indices]
    for i in range(len(v)):
        bits = np.concatenate([v[i]["pos"], v[i]["norm"], v[i]["uv"]]).astype("<f4").view("<u4")
        hv.update(struct.pack("<8Ii", *[int(x) for x in bits], int(mat_of_index[i])))
    for m in mesh.materials:
        f = np.concatenate([m["ambient"][:3], m["diffuse"][:3], m["specular"][:3], m["transmittance"][:3], m["emission"],
                            [m["shininess"], m["ior"], m["dissolve"]]]).astype("<f4").view("<u4")
        hm.update(struct.pack("<18Ii", *[int(x) for x in f], int(m["illum"])))
    return hv.hexdigest(), hm.hexdigest(), len(v), len(mesh.materials)


def oracle_digests():
    from oracle import orc
    sc = S.cornell_scene(with_suzanne=True)
    R, W, H = 64, 160, 120
    view, proj = S.reference_camera(W / H)
    r = orc.render_frame(sc, view, proj, R, W, H)
    d = {"scene": "cornell+suzanne", "R": R, "W": W, "H": H,
         "base": hashlib.sha256(r["base"].tobytes()).hexdigest(),
         "tri_id": hashlib.sha256(r["gbuffer"].tri_id.tobytes()).hexdigest(),
         "fragments": int(r["voxel_stats"].fragments), "occupied": int(r["voxel_stats"].occupied),
         "samples": int(r["trace_stats"].samples), "shaded_pixels": int(r["trace_stats"].shaded_pixels)}
    for l in (1, 3, 6):
        for dr in (0, 3, 5):
            d[f"mip_l{l}_d{dr}"] = hashlib.sha256(r["pyramid"].levels[dr][l].tobytes()).hexdigest()
    # the frame depends on libm (log2f/powf/tanf): keep a coarse digest only (mean colour)
    d["frame_mean_rgba"] = [float(x) for x in r["frame"].view(np.uint8).reshape(-1, 4).mean(axis=0)]
    # non-reference diffuse variants (BASELINE.json configs 1 and 5): 5 and 16 cones on the same pyramid / G-buffer
    for n in (5, 16):
        f, st = orc.trace(sc, view, r["gbuffer"], r["pyramid"], orc.default_params(n_diffuse_cones=n))
        d[f"samples_diffuse_{n}"] = int(st.samples_diffuse)
        d[f"frame_mean_rgba_{n}"] = [float(x) for x in f.view(np.uint8).reshape(-1, 4).mean(axis=0)]
    return d


def main():
    os.makedirs(GOLD, exist_ok=True)
    streams = {}
    for obj in ("CornellBox-Glossy.obj", "suzanne.obj"):
        hv, hm, n, k = tinyobj_stream(os.path.join(REF_ASSETS, obj))
        streams[obj] = {"vertex_stream_sha256": hv, "materials_sha256": hm, "n_indices": n, "n_materials": k}
    json.dump(streams, open(os.path.join(GOLD, "tinyobj_streams.json"), "w"), indent=1)
    json.dump(oracle_digests(), open(os.path.join(GOLD, "oracle_regression.json"), "w"), indent=1)
    print(json.dumps(streams, indent=1))


if __name__ == "__main__":
    main()
