"""The CUDA path (through the C ABI) directly against what the reference's shaders produced on a real OpenGL implementation: the golden
outputs Mesa llvmpipe rendered for tests/test_gl_llvmpipe.py (tests/golden/gl_llvmpipe_*.npz; no GL needed at run time).

Only what that driver computes as the GL specification writes it is compared here (its brilinear mip filter is not: the product
implements the specification's linear blend, see the CPU test): the voxelization fragments, the mip chain, and frames without texture
fetches (raster coverage, near-plane clipping, interpolation, depth test, Blinn-Phong, unorm conversion)."""
import numpy as np
import pytest

import test_gl_llvmpipe as T
from oracle import orc
from voxel_cone_tracing_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    d = capi.Device(0)
    yield d
    d.close()


@pytest.mark.parametrize("name", sorted(T.VOXEL_CASES))
def test_cuda_voxel_grid_vs_llvmpipe_fragments(name):
    """llvmpipe's fragment list (reference's voxelize.vert / .geom / .frag up to the image store) folded in list order = the CUDA grid:
    same occupied voxels, same 4-bit sample counts, colour within one step of the 7-bit average on < 1 % of the voxels."""
    g = np.load(T.VOXEL_GOLDEN)
    tri, vox, col = g[name + ":tri"], g[name + ":voxel"].astype(np.int64), g[name + ":colour"]
    sc, res = T.voxel_scene(name)
    p = capi.Pipeline(sc, res, 64, 64, min(7, int(np.log2(res)) + 1))
    p.clear(); p.voxelize()
    got = p.grid.download(0)
    st = p.voxel_stats()
    p.close()
    assert st.fragments == len(tri)
    grid = np.zeros((res, res, res), np.uint32)
    for i in range(len(tri)):
        x, y, z = vox[i]
        grid[z, y, x] = orc.fold(int(grid[z, y, x]), col[i])
    assert np.array_equal(grid != 0, got != 0), "occupancy differs from llvmpipe's"
    assert not ((grid ^ got) & 0x01010101).any(), "a voxel received a different number of fragments than on llvmpipe"
    d = np.abs(grid.view(np.uint8).astype(int) - got.view(np.uint8).astype(int)).reshape(-1, 4).max(axis=1)
    assert d.max() <= 2 and (d > 0).sum() <= 0.01 * st.occupied, (d.max(), (d > 0).sum(), st.occupied)


@pytest.mark.parametrize("name", ["mip_scene_32", "mip_sparse_random_32", "mip_random_16"])
def test_cuda_mip_chain_vs_llvmpipe(dev, name):
    """llvmpipe's mip chain (reference's mipmap.comp run as a fragment shader) vs the CUDA chain: within one LSB, on rounding ties only."""
    g = np.load(T.MIP_GOLDEN)
    base = T.mip_base(name)
    R = base.shape[0]
    levels = int(np.log2(R)) + 1
    grid = capi.Grid(dev, R, levels)
    grid.upload_base(base)
    capi.check(dev.L.vct_mipmap(dev.h, grid.h))
    n = n_off = 0
    for d in range(6):
        for l in range(1, levels):
            got, gl = grid.download(l, d), g[f"{name}:{d}:{l}"]
            diff = np.abs(got.view(np.uint8).astype(int) - gl.view(np.uint8).astype(int))
            assert diff.max() <= 1, (d, l, diff.max())
            n += gl.size; n_off += int((got != gl).sum())
    grid.close()
    assert n_off <= (0.01 if name == "mip_scene_32" else 0.15) * n, (n_off, n)


@pytest.mark.parametrize("name", ["cornell_direct", "inside_direct"])
def test_cuda_frame_without_texture_fetches_vs_llvmpipe(name):
    """Direct light only, no shadow cone: what is left is GL's fixed function + Blinn-Phong.  Same covered pixels as llvmpipe (camera outside
    and inside the box: near-plane clipping), colours within 1/255 (2/255 on < 0.1 %) of llvmpipe's."""
    gl = np.load(T.GOLDEN)[name]
    sc, view, proj, R, W, H, prm = T.case_inputs(name)
    p = capi.Pipeline(sc, R, W, H, 7)
    p.render_frame(view, proj, capi.default_params(enable_diffuse=0, enable_specular=0, enable_shadow=0))
    got = p.target.frame()
    tri = p.target.gbuffer()["tri_id"]
    p.close()
    assert np.array_equal(tri == 0xFFFFFFFF, gl == T.BACKGROUND), "a pixel is covered on one rasteriser and not on the other"
    d = T.channel_diff(got, gl)
    assert d.max() <= 2 and (d > 1).mean() < 0.001 and (d > 0).mean() < 0.02, (d.max(), (d > 1).mean(), (d > 0).mean())   # 1/255 oracle vs llvmpipe + 1/255 CUDA vs oracle
