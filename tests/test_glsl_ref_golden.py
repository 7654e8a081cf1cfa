"""The restated oracle against golden vectors PRODUCED BY THE REFERENCE'S OWN GLSL (tests/golden/glsl_ref_vectors.json, written by
tools/make_glsl_ref_golden.py from oracle/_ref/libvct_glsl_ref.so, i.e. /root/reference/shader/* executed on the CPU over the
reference's GLM).  Unlike tests/test_glsl_ref.py this needs neither the library nor the reference tree: it runs on any box."""
import base64
import hashlib
import json
import os
import zlib

import numpy as np
import pytest

from oracle import orc
from voxel_cone_tracing_b200 import scene as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "glsl_ref_vectors.json")


@pytest.fixture(scope="module")
def gold():
    with open(GOLD) as f:
        return json.load(f)


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def f32(words) -> np.ndarray:
    return np.array(words, np.uint32).view(np.float32)


def test_fold_sequences(gold):
    for seq in gold["fold_sequences"]:
        vals = f32(seq["vals_f32_hex"]).reshape(-1, 4)
        stored = 0
        for v, exp in zip(vals, seq["stored"]):
            stored = orc.fold(stored, v)
            assert stored == exp


def test_axis_selection(gold):
    tris = f32(gold["axis"]["tris_f32_hex"]).reshape(-1, 3, 3)
    for t, exp in zip(tris, gold["axis"]["axis"]):
        if exp >= 0:
            assert orc.select_axis(t[0], t[1], t[2]) == exp


def test_voxel_grids_and_mip_chains(gold):
    for g in gold["grids"]:
        sc = S.cornell_scene(with_suzanne=g["suzanne"], theta=g["theta"])
        base, st = orc.voxelize(sc, g["R"])
        assert st.fragments + st.fragments_oob == g["fragments_executed"]
        assert st.occupied == g["occupied"]
        if "voxels" in g:
            idx = np.array(g["voxels"]["index"]); val = np.array(g["voxels"]["value"], np.uint32)
            exp = np.zeros(g["R"] ** 3, np.uint32); exp[idx] = val
            bad = np.flatnonzero(exp != base.reshape(-1))
            assert bad.size == 0, f"voxel {bad[0]}: oracle {base.reshape(-1)[bad[0]]:#010x}, reference GLSL {exp[bad[0]]:#010x}"
        assert sha(base) == g["base_sha256"]
        pyr = orc.mipmap(base, g["levels"])
        for key, exp in g["mip_sha256"].items():
            d, l = (int(x) for x in key.split("."))
            assert sha(pyr.levels[d][l]) == exp, f"mip level {l} direction {d}"


def test_cones_and_frames(gold):
    sc = S.cornell_scene(with_suzanne=True, theta=0.3)
    base, _ = orc.voxelize(sc, 64)
    pyr = orc.mipmap(base, 7)
    for c in gold["cones"]:
        got, _ = orc.trace_cone(pyr, f32(c["origin"]), f32(c["dir"]), c["aperture"], c["max_dist"])
        exp = f32(c["rgba_f32_hex"])
        # bit-identical on the libm the vectors were made with; another libm may round log2f differently
        assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)) or np.allclose(got, exp, rtol=0, atol=2e-6)
    for fr in gold["frames"]:
        W, H = fr["W"], fr["H"]
        view, proj = S.reference_camera(W / H, **fr["camera"])
        g = orc.gbuffer(sc, view, proj, W, H)
        hit = (g.tri_id != 0xFFFFFFFF)[..., None]
        assert sha(g.tri_id) == fr["tri_id_sha256"] and sha(g.depth) == fr["depth_sha256"]
        assert sha(np.where(hit, g.world_pos, 0).astype(np.float32)) == fr["world_pos_sha256"]
        assert sha(np.where(hit, g.normal, 0).astype(np.float32)) == fr["normal_sha256"]
        frame, _ = orc.trace(sc, view, g, pyr)
        exp = np.frombuffer(zlib.decompress(base64.b64decode(fr["frame_zlib_b64"])), np.uint32).reshape(H, W)
        if sha(frame) != fr["frame_sha256"]:   # (libm: powf / log2f / tanf) -- then within one 8-bit step, nearly everywhere equal
            d = np.abs(frame.view(np.uint8).astype(np.int32) - exp.view(np.uint8).astype(np.int32))
            assert d.max() <= 1 and (frame != exp).mean() < 0.01
