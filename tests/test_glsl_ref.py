"""The oracle against THE REFERENCE'S OWN GLSL executed on the CPU (oracle/glsl_ref/: the shader text, read from /root/reference
where it lies, rewritten syntactically and compiled against the reference's vendored GLM; the fixed-function GL stages are the
written rules both programs share, oracle/vct_fixed_function.h).

Mode "rules" evaluates the built-ins whose precision GLSL leaves open (normalize, round, matrix products, inverse) by the
oracle's rules R5 / R9: the restated oracle must then equal the reference's shader text BIT FOR BIT -- a slip anywhere in the
restatement (an operand order, a constant, a pair table, a quirk "fixed" by accident) fails here.  Mode "glm" keeps GLM's own
built-ins and bounds what that freedom is worth (one step of the 7-bit running average, 1/255 in the frame).

Needs oracle/_ref/libvct_glsl_ref.so (built here by `make -C oracle ref`; travels to the GPU box prebuilt); where neither the
library nor the reference tree exists the tests skip and tests/test_glsl_ref_golden.py still checks the oracle against vectors
this library produced."""
import os
import re

import numpy as np
import pytest

from oracle import glsl_ref as G
from oracle import orc
from voxel_cone_tracing_b200 import scene as S

pytestmark = pytest.mark.skipif(not G.available(), reason="oracle/_ref/libvct_glsl_ref.so not built and no reference tree to build it from")

REF_SHADERS = "/root/reference/shader"


def byte_diff(a, b):
    return np.abs(a.view(np.uint8).astype(np.int32) - b.view(np.uint8).astype(np.int32))


def count_nibble(w):
    return (w & 1) | ((w >> 7) & 2) | ((w >> 14) & 4) | ((w >> 21) & 8)


# --------------------------------------------------------------------------- the translation itself
@pytest.mark.skipif(not os.path.isdir(REF_SHADERS), reason="reference tree not present")
def test_translation_leaves_shader_bodies_untouched():
    """glsl2cpp.py may only touch declarations (qualifiers, blocks, arrays) and literal suffixes: every statement line of every
    function body must survive verbatim (whitespace and the added `f` suffixes aside)"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("glsl2cpp", os.path.join(os.path.dirname(G.__file__), "glsl_ref", "glsl2cpp.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)

    def squeeze(t):
        t = re.sub(r"(\d\.\d*)f\b", r"\1", t)
        return re.sub(r"\s+", "", t)
    checked = 0
    for name in mod.SHADERS:
        src = open(os.path.join(REF_SHADERS, name)).read()
        gen = squeeze(mod.translate(src, name))
        depth, in_func = 0, False
        for line in src.split("\n"):
            code = line.split("//")[0]
            if depth == 0 and re.match(r"^[\w\[\]]+\s+\w+\s*\([^;]*\)\s*\{?\s*$", code):
                in_func = True           # a function definition; everything else at depth 0 is a declaration
            stmt = in_func and depth >= 1 and code.strip() not in ("", "{", "}")
            if stmt and not re.search(r"\b\w+\s+\w+\s*\[\s*\d*\s*\]\s*;", code):   # local array declarations are rewritten (vec4 values[8];)
                assert squeeze(code) in gen, f"{name}: statement changed by the translation: {code.strip()!r}"
                checked += 1
            depth += code.count("{") - code.count("}")
            if depth == 0 and "}" in code:
                in_func = False
    assert checked > 150


# --------------------------------------------------------------------------- pieces
def test_running_average_fold_equals_oracle():
    """imageAtomicRGBA8Avg (voxelize.frag:95-120) against the oracle's sequential fold: 40-step sequences (count wraps at 16)"""
    rng = np.random.default_rng(3)
    n_tie_diffs = 0
    for seq in range(60):
        stored_o = stored_r = 0
        for step in range(40):
            val = rng.random(4).astype(np.float32) if seq % 3 else np.round(rng.random(4) * 8).astype(np.float32) / 8
            g = G.fold(stored_o, val, "glm")     # one step from the same state: round() half away from zero vs ties-to-even
            stored_o = orc.fold(stored_o, val)
            stored_r = G.fold(stored_r, val, "rules")
            assert stored_r == stored_o, (seq, step)
            assert count_nibble(g) == count_nibble(stored_o)
            d = byte_diff(np.array([g & 0xFEFEFEFE], np.uint32), np.array([stored_o & 0xFEFEFEFE], np.uint32)).max()
            assert d in (0, 2)
            n_tie_diffs += int(d != 0)
    assert n_tie_diffs > 0, "the tie rule of round() is implementation-defined and must be visible on eighths"


def test_axis_selection_equals_oracle():
    """voxelize.geom:25-55 incl. the tie cases (strict >, ties fall through to the (x,z) projection)"""
    rng = np.random.default_rng(5)
    tris = [rng.standard_normal((3, 3)).astype(np.float32) for _ in range(300)]
    # exact ties: normals along (1,1,0), (1,0,1), (0,1,1), (1,1,1)
    tris += [np.array([[0.11, 0.23, 0.37], [0.11 + a, 0.23 + b, 0.37 + c], [0.11 + d, 0.23 + e, 0.37 + f]], np.float32)
             for (a, b, c, d, e, f) in [(1, -1, 0.5, 0.25, -0.25, 1), (1, 0.5, -1, 0.25, 1, -0.25), (0.5, 1, -1, 1, 0.25, -0.25), (1, -1, 0, 0, 1, -1)]]
    for t in tris:
        exp = orc.select_axis(t[0], t[1], t[2])
        for mode in G.MODES:
            got = G.select_axis(t[0], t[1], t[2], mode)
            assert got == exp or got == -2, (t, got, exp)


def test_trace_cone_equals_oracle():
    """trace_cone + sample_voxel + textureLod (voxel_cone_tracing.frag:71-119) on the reference scene's pyramid: float results"""
    sc = S.cornell_scene(with_suzanne=True)
    base, _ = orc.voxelize(sc, 64)
    pyr = orc.mipmap(base, 7)
    rng = np.random.default_rng(9)
    for i in range(400):
        o = rng.random(3).astype(np.float32)
        d = rng.standard_normal(3).astype(np.float32)
        ap = [0.55785173935, 0.1, 0.0174533, 1.2][i % 4]
        md = [1.73205080757, 0.7][i % 2]
        exp, _ = orc.trace_cone(pyr, o, d, ap, md)
        got = G.trace_cone(pyr, o, d, ap, md, "rules")
        assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), (i, got, exp)
        assert np.allclose(G.trace_cone(pyr, o, d, ap, md, "glm"), exp, atol=2e-5)


# --------------------------------------------------------------------------- stages
@pytest.mark.parametrize("R,suzanne,theta", [(128, False, 0.0), (64, True, 0.0), (64, True, 1.1), (96, True, 2.9), (32, True, 0.4)])
def test_voxel_grid_and_mip_chain_bit_exact(R, suzanne, theta):
    sc = S.cornell_scene(with_suzanne=suzanne, theta=theta)
    base, st = orc.voxelize(sc, R)
    levels = 7 if R >= 64 else 6
    pyr = orc.mipmap(base, levels)
    tex, n = G.voxelize(sc, R, "rules")
    assert n == st.fragments + st.fragments_oob
    for i in range(6):
        assert np.array_equal(tex[i], base), f"texture {i}: {(tex[i] != base).sum()} voxels differ from the oracle"
    got = G.mipmap(base, levels, "rules")
    for d in range(6):
        for l in range(1, levels):
            assert np.array_equal(got.levels[d][l], pyr.levels[d][l]), (d, l)
    # GLM's own built-ins: same occupancy and counts, colour within one step of the 7-bit average, on a few voxels only
    texg, _ = G.voxelize(sc, R, "glm")
    assert np.array_equal(texg[0] != 0, base != 0)
    assert np.array_equal(count_nibble(texg[0]), count_nibble(base))
    d = byte_diff(texg[0] & 0xFEFEFEFE, base & 0xFEFEFEFE)
    assert d.max() <= 2 and (texg[0] != base).sum() <= 0.05 * max((base != 0).sum(), 1)
    gotg = G.mipmap(base, levels, "glm")
    assert all(np.array_equal(gotg.levels[d][l], pyr.levels[d][l]) for d in range(6) for l in range(1, levels))


def test_mip_chain_random_grid_bit_exact():
    rng = np.random.default_rng(1)
    R = 64
    base = rng.integers(0, 2 ** 32, (R, R, R), dtype=np.uint64).astype(np.uint32)
    base[rng.random((R, R, R)) > 0.3] = 0
    pyr = orc.mipmap(base, 7)
    for mode in G.MODES:
        got = G.mipmap(base, 7, mode)
        for d in range(6):
            for l in range(1, 7):
                assert np.array_equal(got.levels[d][l], pyr.levels[d][l]), (mode, d, l)


CAMERAS = [dict(), dict(eye=(0.6, 1.3, 2.2), pitch=-12.0, yaw=-105.0), dict(eye=(0.2, 0.9, 0.6), pitch=-10.0, yaw=-100.0),   # the last two: inside the box
           dict(eye=(0.3, 0.2, 0.2), pitch=-20.0, yaw=-60.0)]


@pytest.mark.parametrize("camera", CAMERAS)
def test_gbuffer_bit_exact(camera):
    sc = S.cornell_scene(with_suzanne=True, theta=0.7)
    W, H = 320, 200
    view, proj = S.reference_camera(W / H, **camera)
    exp = orc.gbuffer(sc, view, proj, W, H)
    got = G.gbuffer(sc, view, proj, W, H, "rules")
    assert np.array_equal(got.tri_id, exp.tri_id)
    hit = exp.tri_id != 0xFFFFFFFF
    assert hit.mean() > 0.3
    assert np.array_equal(got.depth, exp.depth) and np.array_equal(got.material[hit], exp.material[hit])
    assert np.array_equal(got.world_pos[hit].view(np.uint32), exp.world_pos[hit].view(np.uint32))
    assert np.array_equal(got.normal[hit].view(np.uint32), exp.normal[hit].view(np.uint32))
    glm = G.gbuffer(sc, view, proj, W, H, "glm")
    assert (glm.tri_id != exp.tri_id).mean() < 2e-3           # an edge pixel may flip when a vertex moves by one ulp
    same = (glm.tri_id == exp.tri_id) & hit
    # (a clip coordinate that moves by one ulp can move a vertex to the next 1/256-pixel snap position: attributes shift by that much)
    assert np.abs(glm.world_pos[same] - exp.world_pos[same]).max() < 5e-4 and np.abs(glm.normal[same] - exp.normal[same]).max() < 5e-3


def test_frame_config1_bit_exact():
    """BASELINE config 1 (CornellBox-Glossy, 128^3, 512x512), whole frame: Renderer::render() from the reference's GLSL vs the oracle"""
    sc = S.cornell_scene()
    view, proj = S.reference_camera(1.0)
    ref = orc.render_frame(sc, view, proj, 128, 512, 512)
    got = G.render_frame(sc, view, proj, 128, 512, 512, mode="rules")
    assert got["fragments"] == ref["voxel_stats"].fragments + ref["voxel_stats"].fragments_oob
    assert np.array_equal(got["base"], ref["base"])
    assert np.array_equal(got["frame"], ref["frame"]), f"{(got['frame'] != ref['frame']).sum()} pixels differ"
    glm = G.render_frame(sc, view, proj, 128, 512, 512, mode="glm")
    assert byte_diff(glm["frame"], ref["frame"]).max() <= 2
    assert (glm["frame"] != ref["frame"]).mean() < 0.02


def test_frame_config2_full_size_bit_exact():
    """BASELINE config 2, the benchmark workload, at full size (256^3, 1920x1080): all four stages from the reference's GLSL"""
    sc = S.cornell_scene()
    view, proj = S.reference_camera(1920 / 1080)
    ref = orc.render_frame(sc, view, proj, 256, 1920, 1080)
    got = G.render_frame(sc, view, proj, 256, 1920, 1080, mode="rules")
    assert np.array_equal(got["base"], ref["base"])
    assert all(np.array_equal(got["pyramid"].levels[d][l], ref["pyramid"].levels[d][l]) for d in range(6) for l in range(1, 7))
    assert np.array_equal(got["gbuffer"].tri_id, ref["gbuffer"].tri_id) and np.array_equal(got["gbuffer"].depth, ref["gbuffer"].depth)
    assert np.array_equal(got["frame"], ref["frame"]), f"{(got['frame'] != ref['frame']).sum()} pixels differ"
    assert ref["trace_stats"].shaded_pixels > 700_000


@pytest.mark.parametrize("kw", [dict(), dict(enable_shadow=0), dict(enable_diffuse=0, enable_specular=0), dict(enable_direct=0),
                                dict(view_voxel_dir=1, view_voxel_lod=1.5), dict(view_voxel_dir=4, view_voxel_lod=0.0), dict(view_voxel_dir=3, view_voxel_lod=3.25)])
@pytest.mark.parametrize("camera", [CAMERAS[0], CAMERAS[2]])
def test_frame_variants_bit_exact(kw, camera):
    """refraction (Suzanne, illum 4), phase toggles (renderer.cpp:368-371), the voxel debug view (:373-374, blended over the clear
    colour), camera outside and inside the box"""
    sc = S.cornell_scene(with_suzanne=True, theta=0.3)
    R, W, H = 64, 192, 128
    view, proj = S.reference_camera(W / H, **camera)
    ref = orc.render_frame(sc, view, proj, R, W, H, orc.default_params(**kw))
    frame = G.shade(sc, view, ref["gbuffer"], ref["pyramid"], orc.default_params(**kw), mode="rules")
    assert np.array_equal(frame, ref["frame"]), f"{(frame != ref['frame']).sum()} pixels differ, max {byte_diff(frame, ref['frame']).max()}"
    assert (ref["frame"] != 0xFF404026).mean() > 0.3


# --------------------------------------------------------------------------- BASELINE configs 3 and 4 at full size
@pytest.mark.parametrize("frame_no", [0, 21])
def test_config3_full_size(frame_no):
    """BASELINE config 3 (Cornell box + rotating Suzanne, 512^3, 2560x1440): voxel grid (all six textures), the 36 mip volumes and the
    G-buffer in full, the frame on every 8th 32x32 tile"""
    import bench
    cfg = bench.CONFIGS[3]
    sc = bench.build_scene(cfg, frame_no)
    R, W, H = cfg["R"], cfg["W"], cfg["H"]
    base, st = orc.voxelize(sc, R)
    tex, n = G.voxelize(sc, R, "rules")
    assert n == st.fragments + st.fragments_oob
    assert all(np.array_equal(tex[i], base) for i in range(6))
    del tex
    po, pg = orc.mipmap(base, 7), G.mipmap(base, 7, "rules")
    assert all(np.array_equal(pg.levels[d][l], po.levels[d][l]) for d in range(6) for l in range(1, 7))
    view, proj = S.reference_camera(W / H)
    go, gg = orc.gbuffer(sc, view, proj, W, H), G.gbuffer(sc, view, proj, W, H, "rules")
    assert np.array_equal(go.tri_id, gg.tri_id) and np.array_equal(go.depth, gg.depth)
    assert np.array_equal(go.world_pos.view(np.uint32), gg.world_pos.view(np.uint32)) and np.array_equal(go.normal.view(np.uint32), gg.normal.view(np.uint32))
    fo, _ = orc.trace(sc, view, go, po, None, 8, frame_no % 8)
    fg = G.shade(sc, view, go, po, None, 8, frame_no % 8, "rules")
    assert np.array_equal(fo, fg) and (fo != 0).sum() > 400_000


def test_config4_voxel_grid_full_size():
    """BASELINE config 4 (1 M synthetic triangles, 512^3): 8.8 M fragments through voxelize.frag's compare-and-swap loop"""
    import bench
    cfg = bench.CONFIGS[4]
    sc = bench.build_scene(cfg)
    base, st = orc.voxelize(sc, cfg["R"])
    tex, n = G.voxelize(sc, cfg["R"], "rules")
    assert n == st.fragments + st.fragments_oob and st.fragments > 8_000_000
    assert np.array_equal(tex[0], base) and np.array_equal(tex[5], base)


def test_fragment_order_is_the_only_freedom_that_matters():
    """GL guarantees no order between the fragments of a draw call and voxelize.frag's running average depends on it: under other valid
    orders the reference's own shader keeps occupancy and the 4-bit counts but moves a third of the occupied voxels by a few 1/255
    (profiles/r02_fixed_function_sensitivity.md).  The oracle / CUDA path reproduce the rule-order execution (R4) exactly."""
    sc = S.cornell_scene(with_suzanne=True)
    R = 64
    base, _ = orc.voxelize(sc, R)
    assert np.array_equal(G.voxelize_variant(sc, R), base)
    for kw in (dict(order=G.ORDER_REVERSED_IN_TRIANGLE), dict(order=G.ORDER_REVERSED_TRIANGLES), dict(order=G.ORDER_RANDOM, seed=7)):
        b = G.voxelize_variant(sc, R, **kw)
        assert np.array_equal(b != 0, base != 0) and np.array_equal(count_nibble(b), count_nibble(base))
        d = byte_diff(b & 0xFEFEFEFE, base & 0xFEFEFEFE)
        assert 0 < d.max() <= 6
        assert (b != base).sum() > 0.02 * (base != 0).sum()
    a, b = G.voxelize_variant(sc, R, order=G.ORDER_RANDOM, seed=7), G.voxelize_variant(sc, R, order=G.ORDER_RANDOM, seed=7)
    assert np.array_equal(a, b)


# --------------------------------------------------------------------------- edge cases
from edge_scenes import EDGE_KINDS, edge_scene  # noqa: E402


@pytest.mark.parametrize("kind", EDGE_KINDS)
def test_edge_cases_bit_exact(kind):
    sc = edge_scene(kind)
    R, W, H = 32, 96, 64
    view, proj = S.reference_camera(W / H, eye=(0.1, 0.2, 1.6))
    ref = orc.render_frame(sc, view, proj, R, W, H, n_levels=6)
    got = G.render_frame(sc, view, proj, R, W, H, n_levels=6, mode="rules")
    st = ref["voxel_stats"]
    assert got["fragments"] == st.fragments + st.fragments_oob
    assert np.array_equal(got["base"], ref["base"])
    assert all(np.array_equal(got["pyramid"].levels[d][l], ref["pyramid"].levels[d][l]) for d in range(6) for l in range(1, 6))
    assert np.array_equal(got["gbuffer"].tri_id, ref["gbuffer"].tri_id)
    assert np.array_equal(got["frame"], ref["frame"])
    if kind == "stack":
        assert st.wrapped_voxels > 0 and st.max_per_voxel >= 80
    if kind == "outside":
        assert st.fragments_oob > 0
    if kind == "empty":
        assert st.fragments == 0 and (ref["frame"] == 0xFF404026).all()
    if kind == "tir":      # some pixels of the tilted pane are totally reflecting (black: NaN), others refract
        pane = ref["gbuffer"].tri_id < 2
        black = (ref["frame"] & 0xFFFFFF) == 0
        assert 0.05 < black[pane].mean() < 0.95
    if kind == "lights":
        assert len(sc.lights) == 12 and (ref["frame"] != 0xFF404026).mean() > 0.1


from edge_scenes import FUZZ_SEEDS, fuzz_case  # noqa: E402


@pytest.mark.parametrize("big", [False, True])
@pytest.mark.parametrize("seed", FUZZ_SEEDS)
def test_random_scenes_bit_exact(seed, big):
    """seeded random workloads (tests/edge_scenes.py::fuzz_case; big = triangles of the size of the cube: dozens of fragments per voxel):
    oracle == the reference's GLSL in every output.  (An offline run over 600 further seeds and 200 big ones found no difference either.)"""
    sc, R, levels, W, H, cam, kw = fuzz_case(seed, big)
    view, proj = S.reference_camera(W / H, **cam)
    ref = orc.render_frame(sc, view, proj, R, W, H, orc.default_params(**kw), levels)
    got = G.render_frame(sc, view, proj, R, W, H, orc.default_params(**kw), levels, mode="rules")
    assert got["fragments"] == ref["voxel_stats"].fragments + ref["voxel_stats"].fragments_oob
    assert np.array_equal(got["base"], ref["base"])
    assert all(np.array_equal(got["pyramid"].levels[d][l], ref["pyramid"].levels[d][l]) for d in range(6) for l in range(1, levels))
    assert np.array_equal(got["gbuffer"].tri_id, ref["gbuffer"].tri_id) and np.array_equal(got["gbuffer"].depth, ref["gbuffer"].depth)
    assert np.array_equal(got["frame"], ref["frame"]), f"{(got['frame'] != ref['frame']).sum()} pixels differ"


# --------------------------------------------------------------------------- host-side matrices
def test_camera_and_model_matrices_match_the_reference_host_code():
    """the camera the benchmark uses (scene.reference_camera, restated in the oracle and in include/vct/math.h) against the reference's own
    Camera struct (src/camera.h compiled where it lies: calc_front, glm::lookAt, glm::perspective with 45.0f taken as RADIANS) and the
    Suzanne model matrix against glm::translate / rotate / scale as src/main.cpp:369-372 applies them"""
    for aspect in (1.0, 1920 / 1080, 2560 / 1440):
        v, p = S.reference_camera(aspect)
        gv, gp = G.camera((0.0, 0.9, 3.0), 0.0, -90.0, 45.0, aspect, 0.1, 100.0)      # main.cpp:107-108
        assert np.array_equal(v.view(np.uint32), gv.view(np.uint32)) and np.array_equal(p.view(np.uint32), gp.view(np.uint32))
        assert np.array_equal(orc.perspective(45.0, aspect, 0.1, 100.0).view(np.uint32), gp.view(np.uint32))
        assert abs(p[5] - 1.0 / np.tan(22.5)) < 1e-6                      # tan(22.5 rad) = 0.5579: an effective vertical field of view of 58.3 degrees
    for cam in CAMERAS[1:]:
        v, p = S.reference_camera(1.5, **cam)
        gv, gp = G.camera(cam["eye"], cam["pitch"], cam["yaw"], 45.0, 1.5, 0.1, 100.0)
        assert np.allclose(v, gv, rtol=0, atol=5e-7) and np.array_equal(p.view(np.uint32), gp.view(np.uint32))   # (libm vs numpy cos / sin: one ulp)
    for theta in (0.0, 0.05, 0.3, 1.05, 2.9, -0.7):
        assert np.allclose(S.mat_trs((0.0, 1.1, -0.5), theta, 0.3), G.model_trs((0.0, 1.1, -0.5), theta, 0.3), rtol=0, atol=1e-7)
