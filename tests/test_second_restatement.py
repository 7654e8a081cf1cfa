"""The C++ oracle against a second, independently written numpy restatement of the reference's GLSL (tests/glsl_numpy.py) on
the reference scene (kept beside tests/test_glsl_ref.py, which runs the shader text itself: this one shares no fixed-function code
with the oracle): a slip in one
of the two restatements of voxelize.frag:95-161 (V4/V5), mipmap.comp:45-100 (M1) or voxel_cone_tracing.frag:80-119 (C2/C3)
shows up as a difference here.  CPU only."""
import numpy as np
import pytest

import glsl_numpy as G
from oracle import orc
from voxel_cone_tracing_b200 import scene as S


@pytest.mark.parametrize("R,suzanne", [(64, True), (128, False)])
def test_voxelizer_v1_to_v5(R, suzanne):
    sc = S.cornell_scene(with_suzanne=suzanne)
    exp, st = orc.voxelize(sc, R)
    got, n_frag = G.voxelize(sc, R)
    assert n_frag == st.fragments, (n_frag, st.fragments)
    assert np.array_equal(got != 0, exp != 0), f"occupancy differs in {((got != 0) != (exp != 0)).sum()} voxels"
    # colour + count nibble: the two restatements evaluate the lighting in float32 with their own operation order, so the
    # colour byte may differ by one quantisation step of the 7-bit running average (2/255); the count bits must be equal
    assert np.array_equal(got & np.uint32(0x01010101), exp & np.uint32(0x01010101)), "sample count nibble differs"
    a = (got & np.uint32(0xFEFEFEFE)).view(np.uint8).astype(np.int32)
    b = (exp & np.uint32(0xFEFEFEFE)).view(np.uint8).astype(np.int32)
    d = np.abs(a - b)
    assert d.max() <= 2, f"max colour difference {d.max()}"
    assert (d != 0).sum() <= 0.002 * (exp != 0).sum() * 4, f"{(d != 0).sum()} channel values differ"


def test_running_average_fold_sequences():
    """V5 alone: random fragment sequences of length 1..40 on one voxel (count wrap at 16 included), scalar oracle vs array numpy"""
    rng = np.random.default_rng(0)
    for _ in range(200):
        n = int(rng.integers(1, 41))
        vals = rng.random((n, 4)).astype(np.float32)
        s_o, s_n = 0, np.zeros(1, np.uint32)
        for v in vals:
            s_o = orc.fold(s_o, v)
            s_n = G.avg_step(s_n, (v * np.float32(255))[None, :])
            assert int(s_n[0]) == s_o


@pytest.mark.parametrize("kind", ["scene", "random", "opaque"])
def test_mip_chain_m1(kind):
    if kind == "scene":
        base, _ = orc.voxelize(S.cornell_scene(with_suzanne=True), 64)
        levels = 7
    else:
        rng = np.random.default_rng(2)
        base = rng.integers(0, 2 ** 32, (32, 32, 32), dtype=np.uint64).astype(np.uint32)
        if kind == "opaque":
            base |= np.uint32(0xFF000000)
        levels = 6
    pyr = orc.mipmap(base, levels)
    chain = G.mip_chain(base, levels)
    for d in range(6):
        for l in range(1, levels):
            assert np.array_equal(chain[d][l], pyr.levels[d][l]), f"dir {d} level {l}: {(chain[d][l] != pyr.levels[d][l]).sum()} texels differ"


def test_cone_march_c2_c3():
    R = 64
    base, _ = orc.voxelize(S.cornell_scene(with_suzanne=True), R)
    pyr = orc.mipmap(base, 7)
    chain = [[pyr.levels[d][l] for l in range(7)] for d in range(6)]
    rng = np.random.default_rng(5)
    n = 120
    origin = (0.5 + (rng.random((n, 3)) - 0.5) * 0.36).astype(np.float32)       # inside / around the box (the box spans ~1/3 of the grid)
    direction = rng.normal(size=(n, 3)).astype(np.float32)
    aperture = rng.choice(np.array([0.55785173935, 0.1, 0.0174533, 0.25], np.float32), n)
    max_dist = np.where(rng.random(n) < 0.5, np.float32(1.73205080757), rng.random(n).astype(np.float32) * 0.3 + 0.05).astype(np.float32)
    got, steps = G.trace_cone(chain, R, origin, direction, aperture, max_dist)
    for i in range(n):
        exp, ns = orc.trace_cone(pyr, origin[i], direction[i], float(aperture[i]), float(max_dist[i]))
        assert ns == steps[i], (i, ns, steps[i])
        assert np.allclose(got[i], exp, rtol=2e-5, atol=2e-6), (i, got[i], exp)
    # single textureLod evaluations, incl. positions outside [0,1]^3 (border) and the LOD clamp at both ends
    pos = (rng.random((60, 3)) * 1.2 - 0.1).astype(np.float32)
    lod = (rng.random(60) * 8 - 1).astype(np.float32)
    for d in range(6):
        got = G.texture_lod(chain[d], pos, lod)
        for i in range(len(pos)):
            assert np.allclose(got[i], orc.texture_lod(pyr, d, pos[i], float(lod[i])), rtol=2e-5, atol=2e-6)
