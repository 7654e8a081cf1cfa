"""The reference's visualisation pass on a real OpenGL implementation (TEST INFRASTRUCTURE ONLY).

oracle/_ref/gl/vct_gl_ref (oracle/gl_ref/gl_harness.c) drives Mesa 18.1 llvmpipe -- the software GL that ships inside Nsight Compute,
loaded behind a stand-in libX11 -- through the reference's three passes with the reference's shader text, read from the reference tree
at run time and changed only where this GL 3.3 driver cannot run it (each change syntactic and asserted, see the three
*_shader_dir_for_driver functions):
  visualize()           Renderer::visualize (src/renderer.cpp:355-390): voxel_cone_tracing.vert unmodified, .frag with its dynamic
                        sampler-array index expanded to a six-way select; voxel textures from a given grid + mip chain
  voxelize_fragments()  Renderer::voxelize (:316-353): voxelize.vert / .geom unmodified, .frag up to the image store (no image
                        load/store in the driver: voxel coordinate and colour of every fragment go to two render targets)
  mip_chain()           Renderer::filter (:283-314): mipmap.comp as a fragment shader (no compute shaders in the driver)
  render_frame()        the three chained

What this pins: the fixed-function half the CPU restatements only write down as rules -- where a triangle's fragments fall in both
rasterising passes, clipping, perspective-correct interpolation, the depth test, textureLod's filtering, blending, the unorm
conversions -- plus the shaders as a GLSL compiler executes them.  What it cannot show: the ORDER in which fragments reach the
running average (no image atomics in this driver; GL defines none): that stays a written rule (R4).

    frame_u8, frame_f32 = gl_ref.visualize(scene, view, proj, pyramid, W, H, params)
"""
from __future__ import annotations

import os
import struct
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
BINARY = os.path.join(_HERE, "_ref", "gl", "vct_gl_ref")
MESA_LIBGL = os.environ.get("VCT_MESA_LIBGL", "/opt/nvidia/nsight-compute/2025.2.1/host/linux-desktop-glibc_2_11_3-x64/Mesa/libGL.so.1")
SHADER_DIR = os.environ.get("VCT_REFERENCE_SHADERS", "/root/reference/shader")


# Mesa 18.1's llvmpipe advertises GL 3.3: indexing a sampler array with a value that is not a constant is undefined there, and this
# build samples tex3D[0] whatever the index (measured: the voxel debug view shows texture 0 for every view_voxel_dir).  The
# reference's fragment shader does it three times (voxel_cone_tracing.frag:83-85,255), so for THIS driver the text gets one
# syntactic change: `textureLod(tex3D[i], p, lod)` becomes a call of a six-way select over constant indices.  Nothing else is touched.
_SELECT = """
vec4 vct_tex3D_select(int i, vec3 p, float lod)
{
  if (i == 0) return textureLod(tex3D[0], p, lod);
  if (i == 1) return textureLod(tex3D[1], p, lod);
  if (i == 2) return textureLod(tex3D[2], p, lod);
  if (i == 3) return textureLod(tex3D[3], p, lod);
  if (i == 4) return textureLod(tex3D[4], p, lod);
  return textureLod(tex3D[5], p, lod);
}
"""


def shader_dir_for_driver(tmp: str, expand_sampler_index: bool = True) -> str:
    """Copies the two shaders into `tmp`, the fragment shader with its dynamic sampler indices expanded (see above)."""
    import re
    for name in ("voxel_cone_tracing.vert", "voxel_cone_tracing.frag"):
        text = open(os.path.join(SHADER_DIR, name)).read()
        if name.endswith(".frag") and expand_sampler_index:
            decl = "uniform sampler3D tex3D[6];"
            assert text.count(decl) == 1
            text, n = re.subn(r"textureLod\(tex3D\[([^\]]+)\]\s*,", r"vct_tex3D_select(int(\1),", text)
            assert n == 4, n
            text = text.replace(decl, decl + "\n" + _SELECT)
        with open(os.path.join(tmp, name), "w") as f:
            f.write(text)
    return tmp


def voxelize_shader_dir_for_driver(tmp: str) -> str:
    """voxelize.vert and voxelize.geom unmodified; voxelize.frag with its image store turned into colour outputs.  llvmpipe 18.1 has no
    image load/store (GL 4.2), so the fragment shader cannot run as written; what it computes up to the store -- the voxel coordinate and
    the colour handed to imageAtomicRGBA8Avg (voxelize.frag:122-161) -- is kept character for character and written to two render targets
    instead.  The compare-and-swap loop itself (voxelize.frag:95-120) is pinned elsewhere (oracle/glsl_ref: the reference's text compiled
    for the CPU)."""
    for name in ("voxelize.vert", "voxelize.geom"):
        with open(os.path.join(tmp, name), "w") as f:
            f.write(open(os.path.join(SHADER_DIR, name)).read())
    text = open(os.path.join(SHADER_DIR, "voxelize.frag")).read()

    def once(old, new):
        nonlocal text
        assert text.count(old) == 1, old
        text = text.replace(old, new)

    once("uniform layout (binding = 2, r32ui) uimage3D tex3D[6];",
         "uniform int vct_grid_res;\nlayout (location = 0) out vec4 vct_voxel;\nlayout (location = 1) out vec4 vct_color;")
    a, b = text.index("void imageAtomicRGBA8Avg("), text.index("void main()")
    text = text[:a] + text[b:]                                                    # the CAS loop needs image atomics
    once("ivec3 dim = imageSize(tex3D[0]);", "ivec3 dim = ivec3(vct_grid_res);")
    once("  for (int i = 0; i < 6; i++)\n    imageAtomicRGBA8Avg(tex3D[i], voxel_pos, final_color);",
         "  vct_voxel = vec4(voxel_pos, 1.0);\n  vct_color = final_color;")
    with open(os.path.join(tmp, "voxelize.frag"), "w") as f:
        f.write(text)
    return tmp


def voxelize_fragments(scene, R: int):
    """The fragments of Renderer::voxelize on llvmpipe, in draw / triangle / row / column order:
    (triangle sequence u32[n], pixel xy u32[n, 2], voxel coordinate f32[n, 3], colour f32[n, 4])."""
    with tempfile.TemporaryDirectory() as d:
        job, out = os.path.join(d, "job.bin"), os.path.join(d, "out.bin")
        with open(job, "wb") as f:
            f.write(b"VCTGLJOB")
            f.write(struct.pack("<9I", 2 * R, 2 * R, R, 0, len(scene.verts), len(scene.indices), len(scene.draws), len(scene.materials), len(scene.lights)))
            f.write(struct.pack("<5i", 1, 1, 1, 1, 7))
            f.write(struct.pack("<2f", 0.0, scene.cube_size))
            f.write(np.eye(4, dtype="<f4").tobytes())
            f.write(np.eye(4, dtype="<f4").tobytes())
            f.write(np.ascontiguousarray(scene.lights).tobytes())
            f.write(np.ascontiguousarray(scene.materials).tobytes())
            f.write(np.ascontiguousarray(scene.draws).tobytes())
            f.write(np.ascontiguousarray(scene.verts).tobytes())
            f.write(np.ascontiguousarray(scene.indices, "<u4").tobytes())
        r = subprocess.run([BINARY, voxelize_shader_dir_for_driver(d), job, out], capture_output=True, text=True,
                           env=dict(os.environ, VCT_MESA_LIBGL=MESA_LIBGL))
        if r.returncode:
            raise RuntimeError(f"vct_gl_ref failed ({r.returncode}): {r.stderr[-2000:]}")
        blob = open(out, "rb").read()
    voxelize_fragments.last_log = r.stderr
    n = struct.unpack_from("<I", blob, 0)[0]
    rec = np.frombuffer(blob, "<u4", n * 10, 4).reshape(n, 10)
    return rec[:, 0].copy(), rec[:, 1:3].copy(), rec[:, 3:6].copy().view("<f4"), rec[:, 6:10].copy().view("<f4")


def mip_shader_dir_for_driver(tmp: str) -> str:
    """mipmap.comp as a fragment shader (llvmpipe 18.1 has no compute shaders): the work-group layout line goes, the image declaration
    becomes six colour outputs + a layer uniform, `gl_GlobalInvocationID` becomes (gl_FragCoord.xy, layer), and each
    `imageStore(dest_tex3D[k], write_pos, value)` becomes `vct_out[k] = (value)`.  fetch_texels, alpha_blend, the offset table and the
    six filter expressions (mipmap.comp:10-100) stay character for character."""
    import re
    text = open(os.path.join(SHADER_DIR, "mipmap.comp")).read()

    def once(old, new):
        nonlocal text
        assert text.count(old) == 1, old
        text = text.replace(old, new)

    once("layout (local_size_x = 8, local_size_y = 8, local_size_z = 8) in;", "")
    once("uniform layout (binding = 0, RGBA8) image3D dest_tex3D[6];", "uniform int vct_layer;\nlayout (location = 0) out vec4 vct_out[6];")
    assert text.count("gl_GlobalInvocationID") == 4
    text = text.replace("gl_GlobalInvocationID", "uvec3(uvec2(gl_FragCoord.xy), uint(vct_layer))")
    text, n = re.subn(r"imageStore\(dest_tex3D\[(\d)\], write_pos, ", r"vct_out[\1] = (", text)
    assert n == 6
    with open(os.path.join(tmp, "mipmap.comp"), "w") as f:
        f.write(text)
    return tmp


def mip_chain(base: np.ndarray, levels: int = 7):
    """Renderer::filter on llvmpipe: [dir][level] -> uint32[N, N, N] for levels 1..levels-1 (index 0 = `base`)."""
    R = base.shape[0]
    with tempfile.TemporaryDirectory() as d:
        job, out = os.path.join(d, "job.bin"), os.path.join(d, "out.bin")
        with open(job, "wb") as f:
            f.write(b"VCTGLMIP")
            f.write(struct.pack("<2I", R, levels))
            f.write(np.ascontiguousarray(base, "<u4").tobytes())
        r = subprocess.run([BINARY, mip_shader_dir_for_driver(d), job, out], capture_output=True, text=True, env=dict(os.environ, VCT_MESA_LIBGL=MESA_LIBGL))
        if r.returncode:
            raise RuntimeError(f"vct_gl_ref failed ({r.returncode}): {r.stderr[-2000:]}")
        blob = open(out, "rb").read()
    res, off = [], 0
    for _dir in range(6):
        chain = [base]
        for l in range(1, levels):
            N = R >> l
            chain.append(np.frombuffer(blob, "<u4", N ** 3, off).reshape(N, N, N).copy())
            off += 4 * N ** 3
        res.append(chain)
    return res


def render_frame(scene, view, proj, R: int, W: int, H: int, params=None, levels: int = 7):
    """Renderer::render on llvmpipe, all three passes chained: its fragments folded by the oracle's restatement of imageAtomicRGBA8Avg in list
    order (rule R4: the one step with no GL behind it), its mip chain of that grid, its frame from those textures.
    Returns dict(frame=uint32[H, W], base=uint32[R, R, R], pyramid=orc.Pyramid, fragments=n)."""
    from . import orc
    tri, xy, vox, col = voxelize_fragments(scene, R)
    base = np.zeros((R, R, R), np.uint32)
    v = vox.astype(np.int64)
    for i in range(len(tri)):
        x, y, z = v[i]
        if 0 <= x < R and 0 <= y < R and 0 <= z < R:      # an image store outside the texture is dropped
            base[z, y, x] = orc.fold(int(base[z, y, x]), col[i])
    chain = mip_chain(base, levels)
    pyr = orc.Pyramid(base, levels)
    for d in range(6):
        for l in range(1, levels):
            pyr.levels[d][l][...] = chain[d][l]
    frame, _ = visualize(scene, view, proj, pyr, W, H, params)
    return dict(frame=frame, base=base, pyramid=pyr, fragments=len(tri))


def available() -> bool:
    return os.path.exists(BINARY) and os.path.exists(MESA_LIBGL) and os.path.exists(os.path.join(SHADER_DIR, "voxel_cone_tracing.frag"))


def visualize(scene, view, proj, pyramid, W: int, H: int, params=None, threads: int | None = None, expand_sampler_index: bool = True):
    """(RGBA8 frame as uint32[H, W], RGBA32F frame as float32[H, W, 4]); row 0 = bottom, like glReadPixels and the library's frame."""
    from . import orc
    params = params or orc.default_params()
    R, levels = pyramid.R, pyramid.n_levels
    with tempfile.TemporaryDirectory() as d:
        job, out = os.path.join(d, "job.bin"), os.path.join(d, "out.bin")
        with open(job, "wb") as f:
            f.write(b"VCTGLJOB")
            f.write(struct.pack("<9I", W, H, R, levels, len(scene.verts), len(scene.indices), len(scene.draws), len(scene.materials), len(scene.lights)))
            f.write(struct.pack("<5i", params.enable_direct, params.enable_diffuse, params.enable_specular, params.enable_shadow, params.view_voxel_dir))
            f.write(struct.pack("<2f", params.view_voxel_lod, scene.cube_size))
            f.write(np.asarray(view, "<f4").reshape(16).tobytes())
            f.write(np.asarray(proj, "<f4").reshape(16).tobytes())
            f.write(np.ascontiguousarray(scene.lights).tobytes())
            f.write(np.ascontiguousarray(scene.materials).tobytes())
            f.write(np.ascontiguousarray(scene.draws).tobytes())
            f.write(np.ascontiguousarray(scene.verts).tobytes())
            f.write(np.ascontiguousarray(scene.indices, "<u4").tobytes())
            f.write(np.ascontiguousarray(pyramid.base, "<u4").tobytes())
            for dirn in range(6):
                for l in range(1, levels):
                    f.write(np.ascontiguousarray(pyramid.levels[dirn][l], "<u4").tobytes())
        env = dict(os.environ, VCT_MESA_LIBGL=MESA_LIBGL)
        if threads is not None:
            env["LP_NUM_THREADS"] = str(threads)
        r = subprocess.run([BINARY, shader_dir_for_driver(d, expand_sampler_index), job, out], capture_output=True, text=True, env=env)
        if r.returncode:
            raise RuntimeError(f"vct_gl_ref failed ({r.returncode}): {r.stderr[-2000:]}")
        blob = open(out, "rb").read()
    n = W * H
    u8 = np.frombuffer(blob, "<u4", n, 0).reshape(H, W).copy()
    f32 = np.frombuffer(blob, "<f4", n * 4, n * 4).reshape(H, W, 4).copy()
    visualize.last_log = r.stderr
    return u8, f32
