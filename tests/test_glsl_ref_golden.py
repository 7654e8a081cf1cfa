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


def test_translator_on_a_synthetic_shader():
    """oracle/glsl_ref/glsl2cpp.py on a hand-written snippet (no reference tree needed): qualifiers and blocks become plain members,
    arrays become glsl_array, literals get their suffix, non-constant global initialisers move into _init_globals(), defines are
    undefined again -- and statements inside functions stay as they are"""
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "glsl_ref", "glsl2cpp.py")
    spec = importlib.util.spec_from_file_location("glsl2cpp", path)
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    src = """#version 450 core
layout (local_size_x = 8) in;
layout (std140, binding = 1) uniform material
{
  vec3 diffuse;
  float shininess;
};
in VS_OUT
{
  vec4 world_position;
} vs_out[];
uniform layout (binding = 2, r32ui) uimage3D tex3D[6];
uniform int cube_res;
out vec4 final_color;
float voxel_size = 1.0 / cube_res;
#define HALF 0.5
const ivec3 offs[] = ivec3[2]
(
  ivec3(1, 0, 0),
  ivec3(0, 1, 0)
);
vec4[2] both(vec4 v)
{
  vec4 r[2];
  r[0] = v * 0.25 + vec4(HALF);
  r[1] = v / 2;
  return r;
}
void main()
{
  final_color = both(vs_out[0].world_position)[1] * voxel_size;
}
"""
    out = mod.translate(src, "synthetic.frag")
    flat = " ".join(out.split())
    assert "#version" not in out and "layout" not in out and "uniform" not in out
    assert "vec3 diffuse;" in flat and "float shininess;" in flat and "material" not in flat          # block flattened
    assert "struct VS_OUT { vec4 world_position; } vs_out[3];" in flat                                 # interface block, one triangle
    assert "glsl_array<uimage3D, 6> tex3D;" in flat and "int cube_res;" in flat and "vec4 final_color;" in flat
    assert "float voxel_size;" in flat and "void _init_globals() { voxel_size = 1.0f / cube_res; }" in flat
    assert "const glsl_array<ivec3, 2> offs = glsl_array<ivec3, 2>{{ ivec3(1, 0, 0), ivec3(0, 1, 0) }};" in flat
    assert "glsl_array<vec4, 2> both(vec4 v)" in flat and "glsl_array<vec4, 2> r;" in flat
    assert "r[0] = v * 0.25f + vec4(HALF);" in flat and "r[1] = v / 2;" in flat                       # statements untouched but for the suffix
    assert "#define HALF 0.5f" in out and out.rstrip().endswith("#undef HALF")
