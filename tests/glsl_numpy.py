"""A SECOND, independently written restatement of the reference shaders -- numpy float32, written from the GLSL text
(shader/voxelize.vert:24-30, voxelize.geom:25-55, voxelize.frag:66-161, mipmap.comp:10-100, voxel_cone_tracing.frag:71-119)
and the GL rules of SURVEY.md appendix A, NOT from oracle/vct_oracle.cpp.  Test infrastructure only.

Purpose: the C++ oracle is the parity anchor of the CUDA path, and no GL driver in this image can run the reference's GLSL
(this module predates oracle/glsl_ref/, which compiles the shader text itself for the CPU).  A transcription slip in the oracle would be copied faithfully by the kernels and
every "CUDA == oracle" test would still pass.  This module narrows that gap: a different author-pass, a different language,
array-at-a-time instead of fragment-at-a-time, compared against the oracle on the reference scene by
tests/test_second_restatement.py.  It shares only the written-down rules (1/256-pixel snapping, top-left fill rule,
barycentrics from the snapped positions, canonical fragment order, round-half-even) with the oracle, not code."""
from __future__ import annotations

import numpy as np

F = np.float32


# ------------------------------------------------------------------------------------------------ V5: imageAtomicRGBA8Avg
def conv_vec4_to_rgba8(v):          # voxelize.frag:66-71, v: (...,4) float32 -> uint32
    u = v.astype(np.uint32) & np.uint32(0xFF)           # uint(float) truncates
    return (u[..., 3] << np.uint32(24)) | (u[..., 2] << np.uint32(16)) | (u[..., 1] << np.uint32(8)) | u[..., 0]


def conv_rgba8_to_vec4(w):          # :73-78
    w = w.astype(np.uint32)
    return np.stack([(w & np.uint32(0xFF)), (w >> np.uint32(8)) & np.uint32(0xFF), (w >> np.uint32(16)) & np.uint32(0xFF), w >> np.uint32(24)], -1).astype(F)


def enc_nibble(m, n):               # :80-86
    m = m.astype(np.uint32); n = n.astype(np.uint32)
    return ((m & np.uint32(0xFEFEFEFE)) | (n & np.uint32(1)) | ((n & np.uint32(2)) << np.uint32(7)) | ((n & np.uint32(4)) << np.uint32(14))
            | ((n & np.uint32(8)) << np.uint32(21)))


def dec_nibble(m):                  # :88-93
    m = m.astype(np.uint32)
    return ((m & np.uint32(1)) | ((m & np.uint32(0x100)) >> np.uint32(7)) | ((m & np.uint32(0x10000)) >> np.uint32(14))
            | ((m & np.uint32(0x1000000)) >> np.uint32(21)))


def avg_step(stored, val255):
    """one successful CAS of the loop at :95-120 applied to arrays of voxels: stored (n,) uint32, val255 (n,4) float32 = val * 255"""
    first = stored == 0
    new_first = enc_nibble(conv_vec4_to_rgba8(val255), np.ones_like(stored))
    rval = conv_rgba8_to_vec4(stored & np.uint32(0xFEFEFEFE))
    n = dec_nibble(stored)
    rval = rval * n.astype(F)[:, None] + val255
    n1 = n + np.uint32(1)
    rval = rval / n1.astype(F)[:, None]
    rval = np.rint(rval / F(2)) * F(2)                  # GLSL round(): half to even (appendix A)
    new_other = enc_nibble(conv_vec4_to_rgba8(rval), n1)
    return np.where(first, new_first, new_other).astype(np.uint32)


# ------------------------------------------------------------------------------------------------ V1-V4: voxelization
def _mat(m16):
    return np.asarray(m16, F).reshape(4, 4).T           # column-major storage -> M[row, col]


def _normalize(v):
    l = np.sqrt((v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1]) + v[..., 2] * v[..., 2])
    return v / l[..., None]


def _dot(a, b):
    return (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]


def voxelize(scene, R: int):
    """returns (grid uint32 [R,R,R] indexed [z,y,x], number of fragments)"""
    cube = F(scene.cube_size)
    frag_voxel, frag_val = [], []
    lights = scene.lights
    for d in scene.draws:                               # draw order = canonical order, renderer.cpp:242-255
        M = _mat(d["model"])
        N3 = np.linalg.inv(M[:3, :3].astype(np.float64)).T.astype(F)     # mat3(transpose(inverse(model)))
        mat = scene.materials[int(d["material"])]
        idx = scene.indices[int(d["first_index"]):int(d["first_index"]) + int(d["index_count"])].reshape(-1, 3)
        vs = scene.verts[int(d["vertex_base"]) + idx]   # (T,3) records
        P = vs["pos"].astype(F)                          # (T,3,3)
        # voxelize.vert:26: (model * vec4(position, 1)) / cube_size
        wp = np.empty_like(P)
        for r in range(3):
            wp[..., r] = ((M[r, 0] * P[..., 0] + M[r, 1] * P[..., 1]) + M[r, 2] * P[..., 2]) + M[r, 3]
        wp = wp / cube
        nr = vs["norm"].astype(F)
        nn = np.empty_like(nr)
        for r in range(3):
            nn[..., r] = (N3[r, 0] * nr[..., 0] + N3[r, 1] * nr[..., 1]) + N3[r, 2] * nr[..., 2]
        nn = _normalize(nn)                              # voxelize.vert:28
        # voxelize.geom:27-29,39-50
        e1, e2 = wp[:, 1] - wp[:, 0], wp[:, 2] - wp[:, 0]
        c = np.abs(np.stack([e1[:, 1] * e2[:, 2] - e2[:, 1] * e1[:, 2], e1[:, 2] * e2[:, 0] - e2[:, 2] * e1[:, 0], e1[:, 0] * e2[:, 1] - e2[:, 0] * e1[:, 1]], -1))
        zdom = (c[:, 2] > c[:, 0]) & (c[:, 2] > c[:, 1])
        xdom = ~zdom & (c[:, 0] > c[:, 1]) & (c[:, 0] > c[:, 2])
        for t in range(len(idx)):
            a_ax, b_ax = (0, 1) if zdom[t] else ((1, 2) if xdom[t] else (0, 2))
            # viewport 2R x 2R (renderer.cpp:339-340): window = (ndc + 1) * R ; snapped to 1/256 pixel
            X = np.rint((wp[t, :, a_ax] + F(1)) * F(R) * F(256)).astype(np.int64)
            Y = np.rint((wp[t, :, b_ax] + F(1)) * F(R) * F(256)).astype(np.int64)
            area2 = (X[1] - X[0]) * (Y[2] - Y[0]) - (Y[1] - Y[0]) * (X[2] - X[0])
            if area2 == 0:
                continue
            sgn = 1 if area2 > 0 else -1
            i0 = max(int(-(-(X.min() - 128) // 256)), 0); i1 = min(int((X.max() - 128) // 256), 2 * R - 1)
            j0 = max(int(-(-(Y.min() - 128) // 256)), 0); j1 = min(int((Y.max() - 128) // 256), 2 * R - 1)
            if i0 > i1 or j0 > j1:
                continue
            jj, ii = np.meshgrid(np.arange(j0, j1 + 1, dtype=np.int64), np.arange(i0, i1 + 1, dtype=np.int64), indexing="ij")   # row-major = canonical order
            px, py = ii * 256 + 128, jj * 256 + 128
            inside = np.ones(px.shape, bool)
            E = []
            for k in range(3):                           # edge opposite vertex k
                a, b = (k + 1) % 3, (k + 2) % 3
                dx, dy = sgn * (X[b] - X[a]), sgn * (Y[b] - Y[a])
                e = sgn * ((X[b] - X[a]) * (py - Y[a]) - (Y[b] - Y[a]) * (px - X[a]))
                top_left = (dy < 0) or (dy == 0 and dx < 0)
                inside &= (e > 0) | ((e == 0) & top_left)
                E.append(e)
            if not inside.any():
                continue
            fa = F(abs(int(area2)))
            b = [E[k][inside].astype(F) / fa for k in range(3)]

            def lerp(attr):                              # attr (3, C): affine interpolation (w = 1)
                return (b[0][:, None] * attr[0][None, :] + b[1][:, None] * attr[1][None, :]) + b[2][:, None] * attr[2][None, :]

            pos = lerp(wp[t])
            nrm = lerp(nn[t])
            # voxelize.frag:122-153
            color = np.zeros((len(pos), 3), F)
            for L in lights[:10]:
                lp = L["position"].astype(F) / cube
                dv = lp[None, :] - pos
                dist = np.sqrt(_dot(dv, dv))
                dirv = dv / dist[:, None]
                att = F(1) / ((F(1) + F(0) * dist) + (F(1) * dist) * dist)
                cs = np.maximum(_dot(_normalize(nrm), dirv), F(0))
                color = color + ((cs * att)[:, None] * L["color"].astype(F)[None, :]) * F(L["intensity"])
            color = mat["diffuse"][:3].astype(F)[None, :] * color + mat["emission"].astype(F)[None, :]
            tr, alpha = np.ones(3, F), F(1)
            if int(mat["illum"]) in (4, 6, 7, 9):
                tr, alpha = mat["transmittance"][:3].astype(F), F(mat["dissolve"])
            val = np.clip(np.concatenate([tr[None, :] * color, np.full((len(pos), 1), alpha, F)], 1), F(0), F(1))
            v = (F(R) * (F(0.5) * pos + F(0.5))).astype(np.int64)     # ivec3(dim * scale_and_bias(pos)): truncation
            ok = ((v >= 0) & (v < R)).all(1)
            frag_voxel.append(((v[ok, 2] * R + v[ok, 1]) * R + v[ok, 0]))
            frag_val.append(val[ok] * F(255))
    grid = np.zeros(R * R * R, np.uint32)
    if not frag_voxel:
        return grid.reshape(R, R, R), 0
    vox = np.concatenate(frag_voxel); val = np.concatenate(frag_val)
    # the k-th fragment of every voxel is folded in round k (stable sort keeps the canonical order inside a voxel)
    order = np.argsort(vox, kind="stable")
    vox, val = vox[order], val[order]
    start = np.r_[0, np.flatnonzero(np.diff(vox)) + 1]
    rank = np.arange(len(vox)) - np.repeat(start, np.diff(np.r_[start, len(vox)]))
    for k in range(int(rank.max()) + 1):
        sel = rank == k
        grid[vox[sel]] = avg_step(grid[vox[sel]], val[sel])
    return grid.reshape(R, R, R), len(vox)


# ------------------------------------------------------------------------------------------------ M1: mipmap.comp
_OFFS = [(1, 1, 1), (1, 1, 0), (1, 0, 1), (1, 0, 0), (0, 1, 1), (0, 1, 0), (0, 0, 1), (0, 0, 0)]      # mipmap.comp:10-20 (x, y, z)
_PAIRS = [[(0, 4), (1, 5), (2, 6), (3, 7)], [(4, 0), (5, 1), (6, 2), (7, 3)], [(0, 2), (1, 3), (5, 7), (4, 6)],
          [(2, 0), (3, 1), (7, 5), (6, 4)], [(0, 1), (2, 3), (4, 5), (6, 7)], [(1, 0), (3, 2), (5, 4), (7, 6)]]   # :59-98 in the shader's order


def _unorm(w):
    return conv_rgba8_to_vec4(w) / F(255)


def mip_level(src, d: int):
    """one dispatch of mipmap.comp for direction d: src uint32 [N,N,N] ([z,y,x]) -> uint32 [N/2]^3"""
    c = [_unorm(src[oz::2, oy::2, ox::2]) for (ox, oy, oz) in _OFFS]     # texelFetch(block_pos + voxel_offsets[i])
    acc = None
    for (f, b) in _PAIRS[d]:
        v = c[f] + (F(1) - c[f][..., 3:4]) * c[b]                         # alpha_blend, :40-43
        acc = v if acc is None else acc + v
    out = np.rint(np.clip(acc / F(4), F(0), F(1)) * F(255))                # imageStore to RGBA8: round to nearest even
    return conv_vec4_to_rgba8(out)


def mip_chain(base, levels: int):
    """[d][l] like the reference's six textures; level 0 is the same array for every d"""
    out = []
    for d in range(6):
        lv = [base]
        for _ in range(1, levels):
            lv.append(mip_level(lv[-1], d))
        out.append(lv)
    return out


# ------------------------------------------------------------------------------------------------ C2/C3: textureLod, sample_voxel, trace_cone
def _texel(vol, x, y, z):
    n = vol.shape[0]
    ok = (x >= 0) & (x < n) & (y >= 0) & (y < n) & (z >= 0) & (z < n)     # CLAMP_TO_BORDER, border (0,0,0,0)
    w = np.where(ok, vol[np.clip(z, 0, n - 1), np.clip(y, 0, n - 1), np.clip(x, 0, n - 1)], np.uint32(0))
    return _unorm(w)


def _trilinear(vol, pos):
    n = vol.shape[0]
    u = pos * F(n) - F(0.5)
    i0 = np.floor(u)
    a = (u - i0).astype(F)
    i0 = i0.astype(np.int64)
    r = np.zeros((len(pos), 4), F)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                w = (a[:, 0] if dx else F(1) - a[:, 0]) * (a[:, 1] if dy else F(1) - a[:, 1]) * (a[:, 2] if dz else F(1) - a[:, 2])
                r = r + w[:, None] * _texel(vol, i0[:, 0] + dx, i0[:, 1] + dy, i0[:, 2] + dz)
    return r


def texture_lod(chain_d, pos, lod):
    """GL_LINEAR_MIPMAP_LINEAR on one directional texture; pos (n,3) in [0,1]^3, lod (n,)"""
    nl = len(chain_d)
    lod = np.clip(lod, F(0), F(nl - 1)).astype(F)
    l0 = np.floor(lod).astype(np.int64)
    l1 = np.minimum(l0 + 1, nl - 1)
    f = (lod - l0.astype(F)).astype(F)
    out = np.zeros((len(pos), 4), F)
    for l in range(nl):
        m0, m1 = l0 == l, (l1 == l) & (f > 0)
        if m0.any():
            out[m0] += (F(1) - f[m0])[:, None] * _trilinear(chain_d[l], pos[m0])
        if m1.any():
            out[m1] += f[m1][:, None] * _trilinear(chain_d[l], pos[m1])
    return out


def trace_cone(chain, R: int, origin, direction, aperture, max_dist):
    """voxel_cone_tracing.frag:88-119 for a batch of rays; returns (rgba (n,4), steps (n,))"""
    origin = np.asarray(origin, F).reshape(-1, 3); direction = _normalize(np.asarray(direction, F).reshape(-1, 3))
    n = len(origin)
    aperture = np.broadcast_to(np.asarray(aperture, F), (n,)); max_dist = np.broadcast_to(np.asarray(max_dist, F), (n,))
    vs = F(1) / F(R)
    col = np.zeros((n, 4), F)
    dist = np.full(n, F(3) * vs, F)
    steps = np.zeros(n, np.int64)
    ix = np.where(direction[:, 0] < 0, 0, 1); iy = np.where(direction[:, 1] < 0, 2, 3); iz = np.where(direction[:, 2] < 0, 4, 5)
    ad = np.abs(direction)
    while True:
        live = (col[:, 3] < F(1)) & (dist < max_dist)
        if not live.any():
            break
        k = np.flatnonzero(live)
        diam = dist[k] * aperture[k]
        pos = direction[k] * dist[k][:, None] + origin[k]
        lod = np.maximum(np.log2(diam * F(R)), F(0)).astype(F)
        s = np.zeros((len(k), 4), F)
        for axis, sel in ((0, ix), (1, iy), (2, iz)):
            for d in range(6):
                m = sel[k] == d
                if m.any():
                    s[m] += ad[k[m], axis][:, None] * texture_lod(chain[d], pos[m], lod[m])
        col[k] = col[k] + (F(1) - col[k, 3])[:, None] * s
        dist[k] = dist[k] + np.maximum(diam / F(2), vs)
        steps[k] += 1
    return col, steps
