"""GPU parity tests: every stage of the CUDA path (through the C ABI) against the CPU oracle on
the same inputs.  Bars (BASELINE.json north_star): voxel occupancy bit-exact, voxel colour <= 1 LSB
(we require bit-exact: the kernels share the oracle's arithmetic rules), frame max-abs <= 2/255 and
PSNR >= 45 dB.  Oracle = CPU restatement, itself equal bit for bit to the reference's own GLSL run on the CPU (tests/test_glsl_ref.py) and held
against the reference's shaders running on Mesa llvmpipe (tests/test_gl_llvmpipe.py, tests/test_gpu_vs_llvmpipe.py)."""
import numpy as np
import pytest

from oracle import orc
from voxel_cone_tracing_b200 import capi
from voxel_cone_tracing_b200 import scene as S

pytestmark = pytest.mark.gpu

FRAME_MAX_ABS = 2        # of 255
FRAME_MIN_PSNR = 45.0    # dB


def psnr(a: np.ndarray, b: np.ndarray) -> float:
    a = a.view(np.uint8).astype(np.float64); b = b.view(np.uint8).astype(np.float64)
    mse = np.mean((a - b) ** 2)
    return 99.0 if mse == 0 else 10.0 * np.log10(255.0 ** 2 / mse)


def max_abs(a, b) -> int:
    return int(np.max(np.abs(a.view(np.uint8).astype(np.int32) - b.view(np.uint8).astype(np.int32))))


def assert_pyramid_equal(grid: capi.Grid, pyr: orc.Pyramid):
    """every level and direction (levels >= 1 live in the stacked mipmapped array the texture units read; download and download_array are the same read-back)"""
    for l in range(pyr.n_levels):
        for d in range(6):
            got = grid.download(l, d)
            exp = pyr.levels[d][l]
            assert np.array_equal(got, exp), f"level {l} dir {d}: {(got != exp).sum()} texels differ"
            if l >= 1:
                arr = grid.download_array(l, d)
                assert np.array_equal(arr, exp), f"array level {l} dir {d}: {(arr != exp).sum()} texels differ"


@pytest.fixture(scope="module")
def dev():
    d = capi.Device(0)
    yield d
    d.close()


# --------------------------------------------------------------------------- mip
@pytest.mark.parametrize("R,levels,density", [(128, 7, 1.0), (64, 7, 1.0), (64, 7, 0.02), (32, 6, 1.0), (16, 5, 0.5), (8, 4, 1.0), (64, 3, 1.0)])
def test_mip_random_grid_bit_exact(dev, R, levels, density):
    rng = np.random.default_rng(1)
    base = rng.integers(0, 2 ** 32, (R, R, R), dtype=np.uint64).astype(np.uint32)
    if density < 1.0:
        base[rng.random((R, R, R)) > density] = 0
    g = capi.Grid(dev, R, levels)
    g.upload_base(base)
    capi.check(dev.L.vct_mipmap(dev.h, g.h))
    assert_pyramid_equal(g, orc.mipmap(base, levels))
    g.close()


def _expected_occupancy(pyr: orc.Pyramid, level: int):
    """occupancy bit of a texel = a voxel of its level-0 support is non-zero (for level >= 1 the OR of the 8 child bits; a
    superset of "the texel is non-zero in some direction", which is what makes skipping a footprint exact); dilated bit
    (x+1,y+1,z+1) = OR over the 2x2x2 footprint whose low corner is (x,y,z)"""
    base = pyr.levels[0][0] != 0
    n = base.shape[0] >> level
    b = 1 << level
    occ = base.reshape(n, b, n, b, n, b).any(axis=(1, 3, 5))
    nonzero = np.zeros_like(occ)
    for d in range(6):
        nonzero |= pyr.levels[d][level] != 0
    assert not (nonzero & ~occ).any(), "a non-zero texel without an occupancy bit would make the skip inexact"
    pad = np.zeros((n + 2,) * 3, bool)
    pad[1:-1, 1:-1, 1:-1] = occ
    dil = np.zeros((n + 1,) * 3, bool)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                dil |= pad[dz:dz + n + 1, dy:dy + n + 1, dx:dx + n + 1]
    return occ, dil


@pytest.mark.parametrize("R,levels,density", [(128, 7, 0.001), (64, 7, 0.01), (32, 6, 0.05), (16, 5, 0.02), (64, 3, 0.3)])
def test_occupancy_masks_match_pyramid(dev, R, levels, density):
    """the zero-footprint skip of the cone tracer is only exact if the masks are: bit clear => every texel of the footprint is zero"""
    rng = np.random.default_rng(7)
    base = rng.integers(0, 2 ** 32, (R, R, R), dtype=np.uint64).astype(np.uint32)
    base[rng.random((R, R, R)) > density] = 0
    base[R - 1, R - 1, R - 1] = 0x01010101   # corner texel: value that filters to zero at coarser levels
    base[0, 0, 0] = 0xFFFFFFFF
    g = capi.Grid(dev, R, levels)
    g.upload_base(base)
    capi.check(dev.L.vct_mipmap(dev.h, g.h))
    pyr = orc.mipmap(base, levels)
    for l in range(levels):
        occ, dil = _expected_occupancy(pyr, l)
        assert np.array_equal(g.occupancy(l, False), occ), f"level {l} occupancy"
        assert np.array_equal(g.occupancy(l, True), dil), f"level {l} dilated occupancy"
    g.close()


def test_mip_sequence_on_one_grid(dev):
    """the mip stage skips rewriting tiles it knows to be zero: a sequence of different grids on ONE grid object
    (dense -> sparse -> empty -> other sparse -> dense) must give the same pyramids as fresh builds"""
    R, levels = 64, 7
    rng = np.random.default_rng(11)
    g = capi.Grid(dev, R, levels)
    for density in (1.0, 0.002, 0.0, 0.004, 0.002, 1.0):
        base = rng.integers(0, 2 ** 32, (R, R, R), dtype=np.uint64).astype(np.uint32)
        base[rng.random((R, R, R)) >= density] = 0
        g.upload_base(base)
        capi.check(dev.L.vct_mipmap(dev.h, g.h))
        pyr = orc.mipmap(base, levels)
        assert_pyramid_equal(g, pyr)
        for l in range(levels):
            occ, dil = _expected_occupancy(pyr, l)
            assert np.array_equal(g.occupancy(l, False), occ) and np.array_equal(g.occupancy(l, True), dil)
    g.close()


def test_mip_empty_grid(dev):
    g = capi.Grid(dev, 64, 7)
    g.upload_base(np.full((64, 64, 64), 0xFFFFFFFF, np.uint32))
    capi.check(dev.L.vct_mipmap(dev.h, g.h))
    g.clear()
    capi.check(dev.L.vct_mipmap(dev.h, g.h))
    for l in range(7):
        assert not g.download(l, 3).any()
    g.close()


# --------------------------------------------------------------------------- voxelize
@pytest.mark.parametrize("R,suzanne", [(128, False), (128, True), (256, True), (64, False)])
def test_voxelize_bit_exact(R, suzanne):
    sc = S.cornell_scene(with_suzanne=suzanne)
    exp, st = orc.voxelize(sc, R)
    p = capi.Pipeline(sc, R, 64, 64)
    for _ in range(2):  # twice: determinism + arena reuse
        p.clear(); p.voxelize()
        got = p.grid.download(0)
        gst = p.voxel_stats()
        assert gst.fragments == st.fragments and gst.occupied == st.occupied and gst.max_per_voxel == st.max_per_voxel
        assert np.array_equal(got != 0, exp != 0), "occupancy differs"
        assert np.array_equal(got, exp), f"{(got != exp).sum()} voxel colours differ"
    p.mipmap()
    assert_pyramid_equal(p.grid, orc.mipmap(exp, 7))
    p.close()


def test_synthetic_many_small_triangles():
    """BASELINE configs 4/5 in miniature: thousands of sub-voxel / sub-pixel triangles (icospheres) take the in-thread
    small-triangle path of both rasterisers; big wall triangles take the 8x8 item path.  Grid, G-buffer and frame vs the oracle."""
    sc = S.synthetic_scene(60_000, 0x5EED0001)
    R, W, H = 64, 320, 240
    exp, st = orc.voxelize(sc, R)
    assert st.fragments > 50_000 and st.tris_no_frag > 10_000       # mostly sub-pixel triangles; many emit nothing (non-conservative raster)
    view, proj = S.reference_camera(W / H, eye=(0.0, 0.3, 2.6))
    p = capi.Pipeline(sc, R, W, H, reserve=1 << 21)
    p.clear(); p.voxelize()
    gst = p.voxel_stats()
    assert gst.fragments == st.fragments and gst.occupied == st.occupied and gst.max_per_voxel == st.max_per_voxel
    assert np.array_equal(p.grid.download(0), exp)
    eg = orc.gbuffer(sc, view, proj, W, H)
    p.gbuffer(view, proj)
    got = p.target.gbuffer()
    assert np.array_equal(got["tri_id"], eg.tri_id)
    hit = eg.tri_id != 0xFFFFFFFF
    assert np.array_equal(got["world_pos"][hit], eg.world_pos[hit]) and np.array_equal(got["normal"][hit], eg.normal[hit])
    for sampler in SAMPLERS:
        p.render_frame(view, proj, capi.default_params(sampler=sampler))
        ref = orc.render_frame(sc, view, proj, R, W, H)
        _check_frame(p.target.frame(), ref)
    p.close()


def test_voxelize_slabs_equal_full():
    sc = S.cornell_scene(with_suzanne=True)
    R = 128
    p = capi.Pipeline(sc, R, 64, 64)
    p.clear(); p.voxelize()
    full = p.grid.download(0)
    acc = np.zeros_like(full)
    for k in range(4):
        p.clear(); p.voxelize(k * R // 4, (k + 1) * R // 4)
        part = p.grid.download(0)
        exp, _ = orc.voxelize(sc, R, k * R // 4, (k + 1) * R // 4)
        assert np.array_equal(part, exp)
        acc += part
    assert np.array_equal(acc, full)
    p.close()


def test_voxelize_many_fragments_per_voxel_and_wrap():
    """Stack 40 coplanar quads in one voxel column: exercises the 16-sample count wrap (voxelize.frag:95-120)
    and the long-list path of the resolve kernel."""
    b = S.SceneBuilder(1.0)
    m = S.default_material(); m["diffuse"][:3] = (0.3, 0.6, 0.9); m["emission"] = (0.1, 0.0, 0.2)
    mid = b.add_material(m)
    v = np.zeros(4, S.VERTEX)
    v["pos"] = [(-0.2, -0.2, 0.1), (0.2, -0.2, 0.1), (0.2, 0.2, 0.1), (-0.2, 0.2, 0.1)]
    v["norm"] = (0, 0, 1)
    mesh = S.Mesh(v, np.array([0, 1, 2, 0, 2, 3] * 40, "<u4"), [(0, 240, -1)], np.zeros(0, S.MATERIAL))
    b.add_mesh(mesh, material_override=mid)
    b.add_light((0.0, 0.0, 0.8))
    sc = b.build()
    exp, st = orc.voxelize(sc, 32)
    assert st.wrapped_voxels > 0 and st.max_per_voxel >= 80
    p = capi.Pipeline(sc, 32, 64, 64, levels=6)
    p.clear(); p.voxelize()
    assert np.array_equal(p.grid.download(0), exp)
    assert p.voxel_stats().max_per_voxel == st.max_per_voxel
    p.close()


@pytest.mark.parametrize("scene_kind", ["cornell", "synthetic", "stack"])
def test_voxelize_fixed_point_accumulation_mode(scene_kind):
    """VCT_ACCUM_FIXED_POINT (north_star: "deterministic integer or fixed-point atomic accumulation path"; a non-reference variant): the
    rounded integer mean of a voxel's fragments by 64-bit atomicAdd -- bit-exact against the oracle's restatement of the same variant,
    identical from run to run, within 3/255 of the reference's ordered running average, and switching back to the ordered mode is clean."""
    if scene_kind == "cornell":
        sc, R = S.cornell_scene(with_suzanne=True), 128
    elif scene_kind == "synthetic":
        sc, R = S.synthetic_scene(60_000, 0x5EED0001), 64     # small / mid triangle paths of the setup kernel
    else:
        b = S.SceneBuilder(1.0)
        m = S.default_material(); m["diffuse"][:3] = (0.3, 0.6, 0.9); m["emission"] = (0.1, 0.0, 0.2)
        v = np.zeros(4, S.VERTEX)
        v["pos"] = [(-0.2, -0.2, 0.1), (0.2, -0.2, 0.1), (0.2, 0.2, 0.1), (-0.2, 0.2, 0.1)]
        v["norm"] = (0, 0, 1)
        b.add_mesh(S.Mesh(v, np.array([0, 1, 2, 0, 2, 3] * 40, "<u4"), [(0, 240, -1)], np.zeros(0, S.MATERIAL)), material_override=b.add_material(m))
        b.add_light((0.0, 0.0, 0.8))
        sc, R = b.build(), 32                                  # 80 fragments per voxel: far past the 16-sample wrap of the ordered mode
    exp_fx, st = orc.voxelize(sc, R, accum_mode=orc.ACCUM_FIXED_POINT)
    exp_ord, _ = orc.voxelize(sc, R)
    p = capi.Pipeline(sc, R, 64, 64, levels=6 if R == 32 else 7, reserve=1 << 21)
    p.dev.set_accum_mode(capi.ACCUM_FIXED_POINT)
    for _ in range(3):
        p.clear(); p.voxelize()
        got = p.grid.download(0)
        gst = p.voxel_stats()
        assert (gst.fragments, gst.occupied, gst.max_per_voxel) == (st.fragments, st.occupied, st.max_per_voxel)
        assert np.array_equal(got, exp_fx), f"{(got != exp_fx).sum()} voxels differ from the fixed-point oracle"
    p.mipmap()
    assert_pyramid_equal(p.grid, orc.mipmap(exp_fx, p.grid.levels))
    if scene_kind != "stack":   # (the stack wraps the reference's 4-bit count: the two modes legitimately differ there)
        d = np.abs(exp_fx.view(np.uint8).astype(np.int32) - exp_ord.view(np.uint8).astype(np.int32))
        assert d.max() <= 3, f"fixed-point vs ordered: max {d.max()}/255"
    p.dev.set_accum_mode(capi.ACCUM_ORDERED)
    p.clear(); p.voxelize()
    assert np.array_equal(p.grid.download(0), exp_ord)
    p.close()


@pytest.mark.parametrize("R,levels,suzanne", [(64, 7, True), (128, 8, True), (32, 6, False)])
def test_fp16_grid_full_mip_chain_variant(R, levels, suzanne):
    """BASELINE config 5's storage variant ("fp16 RGBA grid + full mip chain", NOT the reference's RGBA8 / 7 levels): voxels
    and all 6 x levels mip volumes as four halves per texel -- bit-exact against the oracle run in the same variant mode (levels =
    log2(R) + 1 = the full chain down to one texel) -- and the frame within the 2/255 / 45 dB gate."""
    sc = S.cornell_scene(with_suzanne=suzanne)
    W, H = 320, 200
    view, proj = S.reference_camera(W / H)
    ref = orc.render_frame(sc, view, proj, R, W, H, orc.default_params(), levels, orc.FMT_RGBA16F)
    p = capi.Pipeline(sc, R, W, H, levels, fmt=capi.GRID_RGBA16F)
    for _ in range(2):
        p.render_frame(view, proj, capi.default_params(sampler=capi.SAMPLER_TEX))   # (the variant filters in fp32 whatever the sampler says)
        base = p.grid.download_f16(0)
        assert np.array_equal(base, ref["base"]), f"{(base != ref['base']).sum()} fp16 voxels differ"
        for l in range(1, levels):
            for d in range(6):
                got = p.grid.download_f16(l, d)
                assert np.array_equal(got, ref["pyramid"].levels[d][l]), f"fp16 level {l} dir {d}: {(got != ref['pyramid'].levels[d][l]).sum()} texels differ"
        _check_frame(p.target.frame(), ref)
    gst = p.voxel_stats()
    assert gst.fragments == ref["voxel_stats"].fragments and gst.occupied == ref["voxel_stats"].occupied
    cnt = p.trace_count(view, capi.default_params())
    assert abs(cnt.samples - ref["trace_stats"].samples) <= 5e-2 * ref["trace_stats"].samples
    p.close()


def test_voxelize_arena_overflow_is_reported():
    sc = S.cornell_scene()
    p = capi.Pipeline(sc, 128, 64, 64)
    capi.check(p.dev.L.vct_voxelize_reserve(p.dev.h, 1024))
    # a fresh device starts with the default arena; force a tiny one through a new device
    p.close()
    dev = capi.Device(0)
    capi.check(dev.L.vct_voxelize_reserve(dev.h, 1000))
    scn = capi.DeviceScene(dev, sc); g = capi.Grid(dev, 128, 7)
    capi.check(dev.L.vct_voxelize(dev.h, scn.h, g.h, 0, 128))
    st = capi.VoxelStats()
    import ctypes
    assert dev.L.vct_voxelize_stats(dev.h, ctypes.byref(st)) == -4        # VCT_ERR_OVERFLOW, arena grown
    g.clear()
    capi.check(dev.L.vct_voxelize(dev.h, scn.h, g.h, 0, 128))
    capi.check(dev.L.vct_voxelize_stats(dev.h, ctypes.byref(st)))
    exp, _ = orc.voxelize(sc, 128)
    assert np.array_equal(g.download(0), exp)
    g.close(); scn.close(); dev.close()


# --------------------------------------------------------------------------- G-buffer
@pytest.mark.parametrize("W,H,suzanne", [(512, 512, False), (400, 300, True), (333, 217, True)])
def test_gbuffer_matches_oracle(W, H, suzanne):
    sc = S.cornell_scene(with_suzanne=suzanne)
    view, proj = S.reference_camera(W / H)
    exp = orc.gbuffer(sc, view, proj, W, H)
    p = capi.Pipeline(sc, 32, W, H, levels=6)
    p.gbuffer(view, proj)
    got = p.target.gbuffer()
    assert np.array_equal(got["tri_id"], exp.tri_id)
    hit = exp.tri_id != 0xFFFFFFFF
    assert hit.mean() > 0.3
    assert np.array_equal(got["depth"][hit], exp.depth[hit])
    assert np.array_equal(got["material"][hit], exp.material[hit])
    assert np.array_equal(got["world_pos"][hit], exp.world_pos[hit])
    assert np.array_equal(got["normal"][hit], exp.normal[hit])
    p.close()


@pytest.mark.parametrize("camera", [dict(eye=(0.0, 0.8, 0.5)), dict(eye=(0.3, 0.2, 0.2), pitch=-20.0, yaw=-60.0), dict(eye=(-0.6, 1.2, -0.4), pitch=-35.0, yaw=-150.0)])
def test_gbuffer_near_plane_clipping(camera):
    """camera INSIDE the box (the reference app is a fly-through, src/camera.h:25-58): walls, floor and ceiling cross the camera
    plane.  GL clips them against the near plane (oracle rule R2c); dropping them -- what round 1 did -- leaves holes.  Every pixel
    must hit geometry and visibility, depth and the interpolated attributes must equal the oracle bit for bit."""
    sc = S.cornell_scene(with_suzanne=True)
    W, H = 400, 300
    view, proj = S.reference_camera(W / H, **camera)
    exp = orc.gbuffer(sc, view, proj, W, H)
    hit = exp.tri_id != 0xFFFFFFFF
    assert hit.mean() > 0.999, "inside the closed part of the box every pixel sees a surface"
    p = capi.Pipeline(sc, 32, W, H, levels=6)
    p.gbuffer(view, proj)
    got = p.target.gbuffer()
    assert np.array_equal(got["tri_id"], exp.tri_id)
    assert np.array_equal(got["depth"][hit], exp.depth[hit])
    assert np.array_equal(got["material"][hit], exp.material[hit])
    assert np.array_equal(got["world_pos"][hit], exp.world_pos[hit])
    assert np.array_equal(got["normal"][hit], exp.normal[hit])
    p.close()


@pytest.mark.parametrize("sampler", [capi.SAMPLER_FP32, capi.SAMPLER_TEX])
def test_frame_camera_inside_box(sampler):
    got, ref, _ = _frame_pair(S.cornell_scene(with_suzanne=True), 128, 480, 270, sampler=sampler, camera=dict(eye=(0.2, 0.9, 0.6), pitch=-10.0, yaw=-100.0))
    _check_frame(got, ref)


# --------------------------------------------------------------------------- full frame
SAMPLERS = [capi.SAMPLER_FP32, capi.SAMPLER_TEX]   # software fp32 filtering / texture units: both must pass the frame gate


def _frame_pair(sc, R, W, H, params_kw=None, levels=7, sampler=capi.SAMPLER_FP32, camera=None):
    params_kw = params_kw or {}
    view, proj = S.reference_camera(W / H, **(camera or {}))
    ref = orc.render_frame(sc, view, proj, R, W, H, orc.default_params(**params_kw), levels)
    p = capi.Pipeline(sc, R, W, H, levels)
    prm = capi.default_params(sampler=sampler, **params_kw)
    p.render_frame(view, proj, prm)
    got = p.target.frame()
    cnt = p.trace_count(view, prm)
    p.close()
    return got, ref, cnt


def _check_frame(got, ref):
    exp = ref["frame"]
    assert max_abs(got, exp) <= FRAME_MAX_ABS, f"max abs {max_abs(got, exp)}/255"
    assert psnr(got, exp) >= FRAME_MIN_PSNR, f"PSNR {psnr(got, exp):.2f} dB"


@pytest.mark.parametrize("sampler", SAMPLERS)
def test_frame_config1_cornell_128_512(sampler):
    """BASELINE config 1: CornellBox-Glossy, 128^3, 512x512, 9 diffuse + 1 specular + 1 shadow cone."""
    got, ref, cnt = _frame_pair(S.cornell_scene(), 128, 512, 512, sampler=sampler)
    _check_frame(got, ref)
    ts = ref["trace_stats"]
    assert cnt.shaded_pixels == ts.shaded_pixels
    for k in ("samples_diffuse", "samples_shadow", "samples_specular", "samples_refraction"):
        a, b = getattr(cnt, k), getattr(ts, k)
        assert abs(a - b) <= 5e-2 * max(b, 1), (k, a, b)   # alpha == 1.0 rounding may flip a loop exit by one (invisible) sample


@pytest.mark.parametrize("suzanne,R,W,H", [(False, 128, 512, 512), (True, 128, 400, 300)])
def test_cuda_vs_reference_glsl(suzanne, R, W, H):
    """The CUDA path against THE REFERENCE'S OWN GLSL executed on the CPU (oracle/_ref/libvct_glsl_ref.so: shader/*.vert|geom|frag|comp
    translated syntactically and compiled against the reference's GLM, see oracle/glsl_ref/harness.cpp; built in the dev container,
    ships prebuilt).  BASELINE config 1 and the Suzanne scene (refraction): voxel grid, every mip volume and the G-buffer bit for
    bit, the frame inside the 2/255 / 45 dB gate for both samplers."""
    from oracle import glsl_ref as G
    if not G.available():
        pytest.skip("oracle/_ref/libvct_glsl_ref.so not shipped")
    sc = S.cornell_scene(with_suzanne=suzanne)
    view, proj = S.reference_camera(W / H)
    ref = G.render_frame(sc, view, proj, R, W, H, mode="rules")
    p = capi.Pipeline(sc, R, W, H, 7)
    for sampler in SAMPLERS:
        p.render_frame(view, proj, capi.default_params(sampler=sampler))
        _check_frame(p.target.frame(), ref)
    got = p.grid.download(0)
    assert np.array_equal(got, ref["base"]), f"{(got != ref['base']).sum()} voxels differ from the reference's voxelize.frag"
    assert_pyramid_equal(p.grid, ref["pyramid"])
    gb = p.target.gbuffer()
    hit = ref["gbuffer"].tri_id != 0xFFFFFFFF
    assert np.array_equal(gb["tri_id"], ref["gbuffer"].tri_id)
    assert np.array_equal(gb["world_pos"][hit], ref["gbuffer"].world_pos[hit]) and np.array_equal(gb["normal"][hit], ref["gbuffer"].normal[hit])
    p.close()


@pytest.mark.parametrize("sampler", SAMPLERS)
def test_frame_with_suzanne_refraction(sampler):
    got, ref, cnt = _frame_pair(S.cornell_scene(with_suzanne=True), 128, 400, 300, sampler=sampler)
    _check_frame(got, ref)
    assert cnt.samples_refraction > 0


@pytest.mark.parametrize("sampler", SAMPLERS)
@pytest.mark.parametrize("camera", [dict(eye=(0.6, 1.3, 2.2), pitch=-12.0, yaw=-105.0), dict(eye=(-0.5, 0.4, 1.5), pitch=15.0, yaw=-70.0)])
def test_frame_other_cameras(sampler, camera):
    """camera poses other than main.cpp's default (close to walls / the glossy sphere, grazing angles)"""
    got, ref, _ = _frame_pair(S.cornell_scene(with_suzanne=True, theta=0.7), 128, 480, 270, sampler=sampler, camera=camera)
    _check_frame(got, ref)


@pytest.mark.parametrize("kw", [dict(n_diffuse_cones=5), dict(n_diffuse_cones=16), dict(enable_shadow=0), dict(enable_diffuse=0, enable_specular=0),
                                dict(enable_direct=0), dict(view_voxel_dir=1, view_voxel_lod=1.5), dict(view_voxel_dir=4, view_voxel_lod=0.0)])
@pytest.mark.parametrize("sampler", SAMPLERS)
def test_frame_variants(kw, sampler):
    got, ref, _ = _frame_pair(S.cornell_scene(with_suzanne=True), 64, 256, 192, kw, sampler=sampler)
    _check_frame(got, ref)


@pytest.mark.parametrize("sampler", SAMPLERS)
def test_frame_config2_cornell_256_1080p(sampler):
    """BASELINE config 2 (the benchmark workload) at full size."""
    got, ref, cnt = _frame_pair(S.cornell_scene(), 256, 1920, 1080, sampler=sampler)
    _check_frame(got, ref)
    assert abs(cnt.samples - ref["trace_stats"].samples) <= 5e-2 * ref["trace_stats"].samples


def test_cone_kernel_variants_agree():
    """the production march (3: one-level fetches through the nearest-mip texture object, all diffuse cones of a tile in one warp)
    against one warp per cone slot (2), the march that blends two levels in every fetch (1) and the literal loop (0): same frame
    within 1/255, and each inside the oracle gate; both samplers"""
    sc = S.cornell_scene(with_suzanne=True)
    R, W, H = 128, 480, 270
    view, proj = S.reference_camera(W / H)
    ref = orc.render_frame(sc, view, proj, R, W, H, orc.default_params(), 7)
    p = capi.Pipeline(sc, R, W, H)
    for sampler in SAMPLERS:
        prm = capi.default_params(sampler=sampler)
        frames = {}
        for v in ("3", "2", "1", "0"):
            p.dev.debug_set(capi.DEBUG_CONE_VARIANT, int(v))
            p.render_frame(view, proj, prm)
            frames[v] = p.target.frame().copy()
            _check_frame(frames[v], ref)
        assert np.array_equal(frames["3"], frames["2"]) or max_abs(frames["3"], frames["2"]) <= 1
        assert max_abs(frames["2"], frames["1"]) <= 1
        assert max_abs(frames["1"], frames["0"]) <= 1
    p.close()


def test_async_readback_equals_blocking_readback():
    """vct_target_download_frame_async: a moving object, every frame read back through the copy stream while the next frame renders"""
    import torch
    R, W, H = 64, 320, 200
    view, proj = S.reference_camera(W / H)
    p = capi.Pipeline(S.cornell_scene(with_suzanne=True), R, W, H)
    n = 6
    host = [torch.empty((H, W), dtype=torch.int32).pin_memory().numpy().view(np.uint32) for _ in range(n)]
    tickets = []
    for i in range(n):
        p.scene.upload(S.cornell_scene(with_suzanne=True, theta=0.3 * i))
        p.render_frame(view, proj)
        tickets.append(p.target.frame_async(host[i]))
    for tk in tickets:
        p.target.wait(tk)
    with pytest.raises(capi.VctError):
        p.target.wait(n + 1)
    for i in range(n):
        p.scene.upload(S.cornell_scene(with_suzanne=True, theta=0.3 * i))
        p.render_frame(view, proj)
        assert np.array_equal(p.target.frame(), host[i]), f"frame {i}"
    assert not np.array_equal(host[0], host[1])
    p.close()


def test_frame_16_cone_variant_large_frame():
    """BASELINE config 5's 16-cone variant at a frame large enough for the grouped-diffuse kernel (all cones of a tile in one warp)"""
    got, ref, cnt = _frame_pair(S.cornell_scene(with_suzanne=True), 128, 1920, 1080, dict(n_diffuse_cones=16), sampler=capi.SAMPLER_TEX)
    _check_frame(got, ref)
    assert abs(cnt.samples_diffuse - ref["trace_stats"].samples_diffuse) <= 5e-2 * ref["trace_stats"].samples_diffuse


def test_sparse_frame_sequence_matches_fresh_builds():
    """Frame-to-frame bookkeeping (sparse clear of the occupied list, mip build that skips untouched tiles): a moving object, an
    emptied scene, a grid overwritten behind the voxelizer's back (upload) and a slab-wise voxelization, all on ONE grid -- level 0,
    the whole pyramid (records + arrays) and the occupancy masks must equal fresh oracle builds after every step"""
    R, W, H, levels = 64, 64, 48, 7
    view, proj = S.reference_camera(W / H)
    p = capi.Pipeline(S.cornell_scene(with_suzanne=True), R, W, H, levels)
    rng = np.random.default_rng(5)

    def check(base_exp):
        assert np.array_equal(p.grid.download(0), base_exp)
        pyr = orc.mipmap(base_exp, levels)
        assert_pyramid_equal(p.grid, pyr)
        for l in range(levels):
            occ, dil = _expected_occupancy(pyr, l)
            assert np.array_equal(p.grid.occupancy(l, False), occ) and np.array_equal(p.grid.occupancy(l, True), dil), f"occupancy level {l}"

    for step, theta in enumerate((0.0, 0.9, 1.8, 0.9)):
        sc = S.cornell_scene(with_suzanne=True, theta=theta)
        p.scene.upload(sc)
        p.render_frame(view, proj)                      # clear (sparse from the second frame on) + voxelize + mip
        check(orc.voxelize(sc, R)[0])
    # the Cornell box alone: Suzanne's tiles must go back to zero
    sc = S.cornell_scene()
    p.scene.upload(sc)
    p.render_frame(view, proj)
    check(orc.voxelize(sc, R)[0])
    # level 0 written behind the voxelizer's back: dense paths
    base = rng.integers(0, 2 ** 32, (R, R, R), dtype=np.uint64).astype(np.uint32)
    base[rng.random((R, R, R)) >= 0.01] = 0
    p.grid.upload_base(base)
    p.mipmap()
    check(base)
    # back to voxelization: dense clear, then sparse again
    for theta in (0.3, 1.2):
        sc = S.cornell_scene(with_suzanne=True, theta=theta)
        p.scene.upload(sc)
        p.render_frame(view, proj)
        check(orc.voxelize(sc, R)[0])
    # two slabs voxelized one after the other without a clear in between, then a clear + a full voxelization
    p.clear()
    p.voxelize(0, R // 2); p.voxelize(R // 2, R); p.mipmap()
    check(orc.voxelize(sc, R)[0])
    p.clear(); p.clear()
    p.mipmap()
    check(np.zeros((R, R, R), np.uint32))
    p.voxelize(); p.mipmap()
    check(orc.voxelize(sc, R)[0])
    p.close()


def test_two_grids_share_one_device(dev):
    """the occupied-voxel list lives on the device: a grid may clear sparsely only if the list is still its own.  Two grids of
    different resolution voxelized alternately on one device must each match a fresh oracle build every time"""
    sc_a, sc_b = S.cornell_scene(with_suzanne=True, theta=0.2), S.cornell_scene(with_suzanne=True, theta=1.4)
    ds_a, ds_b = capi.DeviceScene(dev, sc_a), capi.DeviceScene(dev, sc_b)
    g1, g2 = capi.Grid(dev, 64, 7), capi.Grid(dev, 32, 6)
    L = dev.L

    def vox(ds, g):
        g.clear()
        capi.check(L.vct_voxelize(dev.h, ds.h, g.h, 0, g.R))
        capi.check(L.vct_mipmap(dev.h, g.h))

    exp = {(id(ds), g.R): orc.voxelize(sc, g.R)[0] for ds, sc in ((ds_a, sc_a), (ds_b, sc_b)) for g in (g1, g2)}
    for ds, g in ((ds_a, g1), (ds_b, g2), (ds_b, g1), (ds_a, g2), (ds_a, g1), (ds_a, g1), (ds_b, g2)):
        vox(ds, g)
        base = exp[(id(ds), g.R)]
        assert np.array_equal(g.download(0), base)
        assert_pyramid_equal(g, orc.mipmap(base, g.levels))
    for o in (g1, g2, ds_a, ds_b):
        o.close()


def test_tile_split_equals_full_frame():
    sc = S.cornell_scene(with_suzanne=True)
    R, W, H = 64, 320, 200
    view, proj = S.reference_camera(W / H)
    p = capi.Pipeline(sc, R, W, H)
    p.render_frame(view, proj)
    full = p.target.frame().copy()
    parts = []
    for r in range(3):
        p.trace(view, capi.default_params(tile_rank=r, tile_nranks=3))
        parts.append(p.target.frame().copy())
    ty, tx = np.meshgrid(np.arange(H) // 32, np.arange(W) // 32, indexing="ij")
    owner = capi.screen_tile_owner(tx, ty, 3)
    merged = np.zeros_like(full)
    for r in range(3):
        # a rank only writes its own tiles; the rest of its buffer still holds the previous content
        merged[owner == r] = parts[r][owner == r]
    assert np.array_equal(merged, full)
    p.close()


def test_frames_in_flight_equal_the_plain_loop():
    """two pipelines (device objects) rendering alternate frames of a moving object, their traces serialised on the shared low-priority
    stream (VCT_DEBUG_TRACE_LOW_PRIORITY), no host synchronisation in between: every frame equals the one-pipeline render"""
    R, W, H, n = 64, 320, 200, 7
    view, proj = S.reference_camera(W / H)
    scenes = [S.cornell_scene(with_suzanne=True, theta=0.2 + 0.35 * k) for k in range(n)]
    one = capi.Pipeline(scenes[0], R, W, H)
    want = []
    for sc in scenes:
        one.scene.upload(sc); one.render_frame(view, proj)
        want.append((one.target.frame().copy(), one.grid.download(0), one.grid.download(2, 4)))
    one.close()
    pipes = [capi.Pipeline(scenes[0], R, W, H) for _ in range(2)]
    for p in pipes:
        p.dev.debug_set(capi.DEBUG_TRACE_LOW_PRIORITY, 1)
    host = [np.zeros((H, W), np.uint32) for _ in range(n)]
    tickets = []
    for k, sc in enumerate(scenes):
        p = pipes[k & 1]
        p.scene.upload(sc); p.render_frame(view, proj)
        tickets.append((p, p.target.frame_async(host[k])))
        if k >= 2:
            q, tk = tickets[k - 2]
            q.target.wait(tk)
    for q, tk in tickets[-2:]:
        q.target.wait(tk)
    for k in range(n):
        assert np.array_equal(host[k], want[k][0]), f"frame {k}"
    for k in (n - 2, n - 1):   # the grids the two pipelines end with
        p = pipes[k & 1]
        assert np.array_equal(p.grid.download(0), want[k][1]) and np.array_equal(p.grid.download(2, 4), want[k][2])
    for p in pipes:
        p.close()


def test_render_is_deterministic():
    sc = S.cornell_scene(with_suzanne=True)
    view, proj = S.reference_camera(4 / 3)
    p = capi.Pipeline(sc, 128, 400, 300)
    frames = []
    for _ in range(3):
        p.render_frame(view, proj)
        frames.append((p.target.frame().copy(), p.grid.download(0), p.grid.download(3, 2)))
    for f in frames[1:]:
        assert all(np.array_equal(a, b) for a, b in zip(f, frames[0]))
    p.close()


# --------------------------------------------------------------------------- texture_3d.h surface
def test_tex3d_clear_and_box_mip(dev):
    import ctypes
    h = ctypes.c_void_p()
    capi.check(dev.L.vct_tex3d_create(dev.h, 16, 8, 4, 3, ctypes.byref(h)))
    col = (ctypes.c_float * 4)(1.0, 0.5, 0.0, 1.0)
    capi.check(dev.L.vct_tex3d_clear(h, col))
    capi.check(dev.L.vct_tex3d_mip(h))
    out = np.zeros((1, 2, 4), np.uint32)
    capi.check(dev.L.vct_tex3d_download(h, 2, out.ctypes.data))
    assert np.all(out == 0xFF0080FF)
    rng = np.random.default_rng(3)
    src = rng.integers(0, 2 ** 32, (4, 8, 16), dtype=np.uint64).astype(np.uint32)
    capi.check(dev.L.vct_tex3d_upload(h, 0, src.ctypes.data))
    capi.check(dev.L.vct_tex3d_mip(h))
    l1 = np.zeros((2, 4, 8), np.uint32)
    capi.check(dev.L.vct_tex3d_download(h, 1, l1.ctypes.data))
    b = src.view(np.uint8).reshape(4, 8, 16, 4).astype(np.uint32)
    exp = (b.reshape(2, 2, 4, 2, 8, 2, 4).sum(axis=(1, 3, 5)) + 4) // 8
    assert np.array_equal(l1.view(np.uint8).reshape(2, 4, 8, 4), exp.astype(np.uint8))
    assert dev.L.vct_tex3d_create(dev.h, 16, 8, 4, 9, ctypes.byref(ctypes.c_void_p())) == -1   # too many levels (glTexStorage3D rule)
    capi.check(dev.L.vct_tex3d_destroy(h))


# --------------------------------------------------------------------------- corners of the path
from edge_scenes import EDGE_KINDS, edge_scene  # noqa: E402


@pytest.mark.parametrize("kind", EDGE_KINDS)
def test_edge_case_scenes(kind):
    """the 16-sample count wrap, geometry outside the cube, degenerate triangles and NaN normals, twelve lights (the shader clamps to ten)
    on a tilted transmissive quad, a scene that produces nothing: the same scenes tests/test_glsl_ref.py runs through the reference's
    own GLSL.  Voxels, mip volumes and visibility bit for bit, the frame inside the gate."""
    sc = edge_scene(kind)
    R, W, H = 32, 96, 64
    view, proj = S.reference_camera(W / H, eye=(0.1, 0.2, 1.6))
    ref = orc.render_frame(sc, view, proj, R, W, H, n_levels=6)
    p = capi.Pipeline(sc, R, W, H, 6)
    for sampler in SAMPLERS:
        p.render_frame(view, proj, capi.default_params(sampler=sampler))
        _check_frame(p.target.frame(), ref)
    assert np.array_equal(p.grid.download(0), ref["base"])
    assert_pyramid_equal(p.grid, ref["pyramid"])
    assert np.array_equal(p.target.gbuffer()["tri_id"], ref["gbuffer"].tri_id)
    gst, st = p.voxel_stats(), ref["voxel_stats"]
    assert (gst.fragments, gst.occupied, gst.max_per_voxel) == (st.fragments, st.occupied, st.max_per_voxel)
    p.close()


from edge_scenes import FUZZ_SEEDS, fuzz_case  # noqa: E402


@pytest.mark.parametrize("big", [False, True])
@pytest.mark.parametrize("seed", FUZZ_SEEDS)
def test_random_scenes(seed, big):
    """seeded random workloads (tests/edge_scenes.py::fuzz_case: triangle soups under general model matrices, opaque / transmissive
    materials with ior on both sides of 1, zero normals, 0..12 lights, cameras inside the geometry, random phase toggles and debug views),
    the same ones tests/test_glsl_ref.py runs through the reference's GLSL: voxels, mip volumes, visibility and interpolated attributes
    bit for bit, the frame inside the gate with the fp32 sampler (the texture-unit sampler's 9-bit weights are checked on the reference's
    own scenes: with random light intensities a shadow term can carry their rounding past 2/255)."""
    if big and seed >= 8:
        pytest.skip("eight seeds of the big-triangle variant (dozens of fragments per voxel: count wrap, long lists in the resolve pass)")
    sc, R, levels, W, H, cam, kw = fuzz_case(seed, big)
    view, proj = S.reference_camera(W / H, **cam)
    ref = orc.render_frame(sc, view, proj, R, W, H, orc.default_params(**kw), levels)
    p = capi.Pipeline(sc, R, W, H, levels)
    p.render_frame(view, proj, capi.default_params(sampler=capi.SAMPLER_FP32, **kw))
    assert np.array_equal(p.grid.download(0), ref["base"])
    assert_pyramid_equal(p.grid, ref["pyramid"])
    gb = p.target.gbuffer()
    hit = ref["gbuffer"].tri_id != 0xFFFFFFFF
    assert np.array_equal(gb["tri_id"], ref["gbuffer"].tri_id)
    for key, exp in (("world_pos", ref["gbuffer"].world_pos[hit]), ("normal", ref["gbuffer"].normal[hit])):
        got = gb[key][hit]
        nan = np.isnan(exp)          # (normals of length zero: the payload / sign of a NaN is not part of the contract)
        assert np.array_equal(np.isnan(got), nan), key
        bad = got.view(np.uint32)[~nan] != exp.view(np.uint32)[~nan]
        assert not bad.any(), f"{key}: {int(bad.sum())} of {bad.size} values differ, max abs {np.abs(got[~nan] - exp[~nan]).max():.3g}"
    _check_frame(p.target.frame(), ref)
    p.render_frame(view, proj, capi.default_params(sampler=capi.SAMPLER_TEX, **kw))
    tex = p.target.frame()
    assert psnr(tex, ref["frame"]) >= 40.0 and (np.abs(tex.view(np.uint8).astype(np.int32) - ref["frame"].view(np.uint8).astype(np.int32)) > 2).mean() < 0.01
    p.close()


@pytest.mark.parametrize("kind", ["degenerate", "tir", "lights"])
def test_edge_case_scenes_large_frame(kind):
    """the same corner cases at 1920x1080: frames of >= 32 k live tiles take the grouped path of the cone kernel (all diffuse cones of a
    tile in one warp, their sum stored), where a NaN cone has to survive the summation"""
    sc = edge_scene(kind)
    R, W, H = 32, 1920, 1080
    view, proj = S.reference_camera(W / H, eye=(0.05, 0.1, 0.9))
    ref = orc.render_frame(sc, view, proj, R, W, H, n_levels=6)
    assert ref["trace_stats"].shaded_pixels > 300_000
    p = capi.Pipeline(sc, R, W, H, 6)
    for sampler in SAMPLERS:
        p.render_frame(view, proj, capi.default_params(sampler=sampler))
        _check_frame(p.target.frame(), ref)
    p.close()
