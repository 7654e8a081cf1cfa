// gbuffer.cu -- camera visibility pass producing the G-buffer the cone tracer shades.
//
// Replaces voxel_cone_tracing.vert + the fixed-function clip / raster / GL_LESS depth test of
// Renderer::visualize() (src/renderer.cpp:355-390).  The reference is a forward renderer that
// shades every overdrawn fragment; the last writer of a pixel is the GL_LESS winner, so shading
// only that fragment gives the same image.  Three launches:
//   cam_setup_kernel    triangle-parallel vertex shader + NEAR-PLANE CLIPPING (a triangle becomes 0, 1 or 2
//                       pieces, rule R2c of the oracle) + projection + snapping.  Pieces whose bounding box
//                       holds a few dozen pixel centres are depth-tested right there (one lane per triangle);
//                       the others get a record in a compact array and 8x8 .. 64x64-pixel work items (size per triangle, raster.cuh).
//   cam_raster_kernel   one warp per work item (8x8 blocks culled by the edge functions, then two pixels per
//                       lane); 64-bit atomicMin of (depth bits << 32 | triangle sequence) = GL_LESS with
//                       "first drawn wins ties"
//   cam_resolve_kernel  per pixel: perspective-correct world position and (un-renormalised) normal of the
//                       winning triangle (voxel_cone_tracing.vert:22-28) -- from its record, or, for the
//                       triangles that never got one (the millions of sub-tile triangles of a large scene),
//                       by running the vertex stage again for that one triangle.
// Built with -fmad=false; same evaluation order as the oracle (rules R1-R3, R2c, R8).
#include "raster.cuh"

namespace vct {

struct Mat4 { float m[16]; };
constexpr int kSmallCamPixels = 100;   // swept together with the voxelizer's limit (tools/small_limit_sweep.py)
// Pieces with up to this many pixel centres in their bounding box (and more than kSmallCamPixels) are rasterised by the whole warp
// inside the set-up kernel, one piece after the other, 8 x 4 pixels per step: no record, no work item, no prefix search.  At 8K with
// 4 M triangles nearly every triangle is of this size (a few dozen to a few hundred pixels); round 1 gave each of them a record + items
// (or, once the record array was full, one lane walking the whole box): 29 + 13 ms of the 33 ms frame share of this pass.
constexpr int kMidCamPixels = 4096;

// one clip-space vertex with the attributes the fragment stage interpolates
struct ClipVert { float cx, cy, cz, cw; float world[3], nn[3]; };

// a + t * (b - a), fp32, no FMA (gbuffer.o is built with -fmad=false): the oracle's clip interpolation
__device__ __forceinline__ float clip_lerp(float a, float b, float t) { return a + t * (b - a); }

// Vertex stage of triangle t (voxel_cone_tracing.vert:22-28): clip-space position, world position, normalised normal per vertex, and the
// signed distance to the near plane.  Static indexing only: everything stays in registers.
__device__ __forceinline__ int cam_vertex_stage(const vct_vertex_t* __restrict__ verts, const uint32_t* __restrict__ indices, const DrawRec& d, uint32_t t,
                                                const float* __restrict__ p, ClipVert (&in)[3], float (&dn)[3]) {
  const uint32_t first = d.first_index + 3u * (t - d.tri_base);
  const float* m = d.model;
  int n_out = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const vct_vertex_t vx = verts[d.vertex_base + indices[first + k]];
    const float px = vx.pos[0], py = vx.pos[1], pz = vx.pos[2];
    const float wx = ((m[0] * px + m[4] * py) + m[8] * pz) + m[12];
    const float wy = ((m[1] * px + m[5] * py) + m[9] * pz) + m[13];
    const float wz = ((m[2] * px + m[6] * py) + m[10] * pz) + m[14];
    const float ww = ((m[3] * px + m[7] * py) + m[11] * pz) + m[15];
    in[k].world[0] = wx; in[k].world[1] = wy; in[k].world[2] = wz;
    const float* nm = d.nmat;
    const float nx = (nm[0] * vx.norm[0] + nm[3] * vx.norm[1]) + nm[6] * vx.norm[2];
    const float ny = (nm[1] * vx.norm[0] + nm[4] * vx.norm[1]) + nm[7] * vx.norm[2];
    const float nz = (nm[2] * vx.norm[0] + nm[5] * vx.norm[1]) + nm[8] * vx.norm[2];
    const float nl = sqrtf((nx * nx + ny * ny) + nz * nz);
    in[k].nn[0] = nx / nl; in[k].nn[1] = ny / nl; in[k].nn[2] = nz / nl;
    in[k].cx = ((p[0] * wx + p[4] * wy) + p[8] * wz) + p[12] * ww;
    in[k].cy = ((p[1] * wx + p[5] * wy) + p[9] * wz) + p[13] * ww;
    in[k].cz = ((p[2] * wx + p[6] * wy) + p[10] * wz) + p[14] * ww;
    in[k].cw = ((p[3] * wx + p[7] * wy) + p[11] * wz) + p[15] * ww;
    dn[k] = in[k].cz + in[k].cw;          // >= 0: inside the near plane (-w <= z)
    if (!(dn[k] >= 0.0f)) n_out++;        // also NaN
  }
  return n_out;
}

// perspective division + viewport + snapping of one (clipped or whole) triangle; false = it produces no fragment
__device__ __forceinline__ bool cam_make_piece(const ClipVert& c0, const ClipVert& c1, const ClipVert& c2, int W, int H, uint32_t material, CamTri& v) {
  float xw[3], yw[3];
  bool ok = true;
  const ClipVert* cs[3] = {&c0, &c1, &c2};
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const ClipVert& c = *cs[k];
#pragma unroll
    for (int a = 0; a < 3; a++) { v.world[k][a] = c.world[a]; v.nn[k][a] = c.nn[a]; }
    if (!(c.cw > 0.0f)) ok = false;   // cannot happen behind a near plane with near > 0; guards general matrices
    const float iw = 1.0f / c.cw;
    v.iw[k] = iw;
    xw[k] = (c.cx * iw + 1.0f) * ((float)W * 0.5f);
    yw[k] = (c.cy * iw + 1.0f) * ((float)H * 0.5f);
    v.zw[k] = (c.cz * iw + 1.0f) * 0.5f;
  }
  v.material = material;
  v.pad = 0;
  return ok && raster_setup(xw, yw, W, H, v.rt);
}

// R2c: Sutherland-Hodgman against the near plane for a triangle that crosses it (1 or 2 vertices outside); a new vertex is always computed
// from the INSIDE vertex of its edge.  Up to two pieces: the fan (p0,p1,p2), (p0,p2,p3) of the clipped polygon.  Rare (the triangles around
// the camera) and deliberately NOT inlined, and it runs the vertex stage again instead of taking its results: dynamically indexed arrays and
// anything whose address crosses a call live in local memory, and in round 1 every one of 4 M triangle set-ups paid for that 608-byte frame.
__device__ __noinline__ int cam_clip_pieces(const vct_vertex_t* __restrict__ verts, const uint32_t* __restrict__ indices, const DrawRec* __restrict__ dp, uint32_t t,
                                            const float* __restrict__ p, int W, int H, CamTri* out) {
  const DrawRec& d = *dp;
  ClipVert in[3];
  float dn[3];
  cam_vertex_stage(verts, indices, d, t, p, in, dn);
  ClipVert poly[4];
  int np = 0;
  for (int k = 0; k < 3; k++) {
    const int k1 = (k + 1) % 3;
    const bool in_a = dn[k] >= 0.0f, in_b = dn[k1] >= 0.0f;
    if (in_a) poly[np++] = in[k];
    if (in_a != in_b) {
      const ClipVert& vi = in_a ? in[k] : in[k1];
      const ClipVert& vo = in_a ? in[k1] : in[k];
      const float di = in_a ? dn[k] : dn[k1], dout = in_a ? dn[k1] : dn[k];
      const float tt = di / (di - dout);
      ClipVert c;
      c.cx = clip_lerp(vi.cx, vo.cx, tt); c.cy = clip_lerp(vi.cy, vo.cy, tt); c.cz = clip_lerp(vi.cz, vo.cz, tt); c.cw = clip_lerp(vi.cw, vo.cw, tt);
      for (int a = 0; a < 3; a++) { c.world[a] = clip_lerp(vi.world[a], vo.world[a], tt); c.nn[a] = clip_lerp(vi.nn[a], vo.nn[a], tt); }
      poly[np++] = c;
    }
  }
  int n_pieces = 0;
  for (int piece = 0; piece + 3 <= np; piece++)
    if (cam_make_piece(poly[0], poly[piece + 1], poly[piece + 2], W, H, d.material, out[n_pieces])) n_pieces++;
  return n_pieces;
}

// tile_rank carries the lattice step in its high half (launch_gbuffer): rank | screen_tile_k(nranks) << 16.  The multi-GPU test is a
// call, not inline: written inline, the compiler evaluates the modulo speculatively inside the rasterisers' pixel loops also when
// tile_nranks == 1 (measured: cam_setup_kernel of the 1 M-triangle scene on one GPU 665 -> 1040 us).
__device__ __noinline__ bool tile_owned_split(int i, int j, int tile_rank, int tile_nranks) {
  return screen_tile_owner(i >> 5, j >> 5, tile_nranks, tile_rank >> 16) == (tile_rank & 0xFFFF);
}
__device__ __forceinline__ bool tile_owned(int i, int j, int W, int tile_rank, int tile_nranks) {
  if (tile_nranks <= 1) return true;
  return tile_owned_split(i, j, tile_rank, tile_nranks);
}
// the set-up kernel is compiled twice (SPLIT = the frame is shared between ranks): its one-GPU build carries no ownership code at all
template <bool SPLIT>
__device__ __forceinline__ bool tile_owned_t(int i, int j, int tile_rank, int tile_nranks) {
  if (!SPLIT) return true;
  return screen_tile_owner(i >> 5, j >> 5, tile_nranks, tile_rank >> 16) == (tile_rank & 0xFFFF);
}

// depth test of every covered pixel of a piece, one lane per piece (the sub-tile triangles of a large scene)
template <bool SPLIT>
__device__ __forceinline__ void raster_piece_inline(const CamTri& v, uint32_t t, int W, unsigned long long* __restrict__ vis, int tile_rank, int tile_nranks) {
  const int bw = v.rt.imax - v.rt.imin + 1, bh = v.rt.jmax - v.rt.jmin + 1;
  EdgeBlock eb;   // edge functions once at the box origin; a pixel then costs two 32 x 32 -> 64 multiply-adds per edge (same integers as raster_sample)
  edge_block_setup(v.rt, v.rt.imin, v.rt.jmin, eb);
  for (int j = v.rt.jmin; j < v.rt.jmin + bh; j++)
    for (int i = v.rt.imin; i < v.rt.imin + bw; i++) {
      if (!tile_owned_t<SPLIT>(i, j, tile_rank, tile_nranks)) continue;   // multi-GPU: not this rank's screen tile
      float b[3];
      if (edge_block_sample(eb, i - v.rt.imin, j - v.rt.jmin, b)) {
        const float zw = interp3(b, v.zw[0], v.zw[1], v.zw[2]);
        if (zw >= 0.0f && zw <= 1.0f) atomicMin(&vis[(size_t)j * W + i], ((unsigned long long)__float_as_uint(zw) << 32) | (unsigned long long)t);
      }
    }
}

// depth test of every covered pixel of ONE piece by all 32 lanes: the owner lane's set-up travels by shuffles, the lanes sweep the
// bounding box in 8 x 4 pixel steps.  Called by the whole warp (src = owner lane); same arithmetic as raster_piece_inline.
template <bool SPLIT>
__device__ __forceinline__ void raster_piece_warp(const CamTri& mine, uint32_t my_tri, int src, int lane, int W, unsigned long long* __restrict__ vis,
                                                  int tile_rank, int tile_nranks) {
  RasterTri rt;
#pragma unroll
  for (int k = 0; k < 3; k++) { rt.X[k] = __shfl_sync(0xffffffffu, mine.rt.X[k], src); rt.Y[k] = __shfl_sync(0xffffffffu, mine.rt.Y[k], src); }
  rt.sign = __shfl_sync(0xffffffffu, mine.rt.sign, src);
  rt.imin = __shfl_sync(0xffffffffu, mine.rt.imin, src); rt.imax = __shfl_sync(0xffffffffu, mine.rt.imax, src);
  rt.jmin = __shfl_sync(0xffffffffu, mine.rt.jmin, src); rt.jmax = __shfl_sync(0xffffffffu, mine.rt.jmax, src);
  rt.mshift = 0;
  rt.area = __shfl_sync(0xffffffffu, mine.rt.area, src);
  const float z0 = __shfl_sync(0xffffffffu, mine.zw[0], src), z1 = __shfl_sync(0xffffffffu, mine.zw[1], src), z2 = __shfl_sync(0xffffffffu, mine.zw[2], src);
  const uint32_t t = __shfl_sync(0xffffffffu, my_tri, src);
  const int lx = lane & 7, ly = lane >> 3;
  EdgeBlock eb;
  edge_block_setup(rt, rt.imin, rt.jmin, eb);
  // multi-GPU: steps aligned to the 8 x 4 grid -- a step then lies inside one 32 x 32 screen tile and "is it this rank's tile" is one test
  // per step.  One GPU: steps start at the corner of the bounding box (a 10 x 10-pixel box is 2 x 3 steps, not up to 3 x 4).
  for (int j0 = SPLIT ? rt.jmin & ~3 : rt.jmin; j0 <= rt.jmax; j0 += 4)
    for (int i0 = SPLIT ? rt.imin & ~7 : rt.imin; i0 <= rt.imax; i0 += 8) {
      if (!tile_owned_t<SPLIT>(i0, j0, tile_rank, tile_nranks)) continue;   // another rank shades (and rasterises) this tile
      const int i = i0 + lx, j = j0 + ly;
      if (i < rt.imin || i > rt.imax || j < rt.jmin || j > rt.jmax) continue;
      float b[3];
      if (edge_block_sample(eb, i - rt.imin, j - rt.jmin, b)) {
        const float zw = interp3(b, z0, z1, z2);
        if (zw >= 0.0f && zw <= 1.0f) atomicMin(&vis[(size_t)j * W + i], ((unsigned long long)__float_as_uint(zw) << 32) | (unsigned long long)t);
      }
    }
}

// does the pixel box [i0,i1] x [j0,j1] touch a 32 x 32 screen tile of this rank?  (boxes of up to 3 x 3 tiles are tested exactly, larger ones kept)
template <bool SPLIT>
__device__ __forceinline__ bool box_touches_owned_tile(int i0, int i1, int j0, int j1, int W, int tile_rank, int tile_nranks) {
  if (!SPLIT) return true;
  const int tx0 = i0 >> 5, tx1 = i1 >> 5, ty0 = j0 >> 5, ty1 = j1 >> 5;
  if (tx1 - tx0 > 2 || ty1 - ty0 > 2) return true;
  for (int ty = ty0; ty <= ty1; ty++)
    for (int tx = tx0; tx <= tx1; tx++)
      if (tile_owned_t<SPLIT>(tx << 5, ty << 5, tile_rank, tile_nranks)) return true;
  return false;
}

// A triangle that crosses the near plane, start to finish (called by its own lane only): clip, cull, then every piece either gets a record
// + work items (returned count) or is depth-tested in line.  Sets big_slot[t].
template <bool SPLIT>
__device__ __noinline__ uint32_t cam_setup_clipped(const vct_vertex_t* __restrict__ verts, const uint32_t* __restrict__ indices, const DrawRec* __restrict__ d, uint32_t t,
                                                   const float* __restrict__ pvm, int W, int H, CamTri* __restrict__ recs, uint32_t rec_capacity,
                                                   uint32_t* __restrict__ rec_count, uint32_t* __restrict__ big_slot, unsigned long long* __restrict__ vis,
                                                   int tile_rank, int tile_nranks, int record_limit) {
  CamTri pc[2];
  const int n = cam_clip_pieces(verts, indices, d, t, pvm, W, H, pc);
  bool big[2] = {false, false};
  int n_big = 0;
  for (int q = 0; q < n; q++) {
    const RasterTri& rt = pc[q].rt;
    if (!box_touches_owned_tile<SPLIT>(rt.imin, rt.imax, rt.jmin, rt.jmax, W, tile_rank, tile_nranks)) { pc[q].rt.sign = 0; continue; }
    big[q] = (rt.imax - rt.imin + 1) * (rt.jmax - rt.jmin + 1) > record_limit;
    n_big += big[q] ? 1 : 0;
  }
  uint32_t slot = 0, count = 0;
  if (n_big) {
    slot = atomicAdd(rec_count, (uint32_t)n_big);
    if (slot + (uint32_t)n_big > rec_capacity) { big[0] = big[1] = false; n_big = 0; }   // record array full: in line (slow, still exact)
  }
  big_slot[t] = n_big ? ((slot + 1u) | ((uint32_t)(n_big - 1) << 31)) : 0u;
  for (int q = 0; q < n; q++) {
    if (pc[q].rt.sign == 0) continue;
    if (big[q]) {
      if (record_limit > 0) raster_choose_macro(pc[q].rt, kMaxItemsManyTris);
      pc[q].pad = raster_item_count(pc[q].rt);
      count += pc[q].pad;
      recs[slot++] = pc[q];
    } else {
      raster_piece_inline<SPLIT>(pc[q], t, W, vis, tile_rank, tile_nranks);
    }
  }
  return count;
}

// big_slot[t]: 0 = the triangle has no records (culled, or rasterised in line: the resolve kernel re-runs the vertex stage);
// else 1 + index of its first record in `recs` | (number of records - 1) << 31.  recs[slot].pad = work items of that record.
template <bool SPLIT>
__global__ void __launch_bounds__(kSetupThreads)
cam_setup_kernel(const vct_vertex_t* __restrict__ verts, const uint32_t* __restrict__ indices, const DrawRec* __restrict__ draws,
                 uint32_t n_draws, uint32_t n_tris, Mat4 pv, int W, int H, CamTri* __restrict__ recs, uint32_t rec_capacity, uint32_t* __restrict__ rec_count,
                 uint32_t* __restrict__ big_slot, uint32_t* __restrict__ item_local, uint32_t* __restrict__ item_block, unsigned long long* __restrict__ vis,
                 int tile_rank, int tile_nranks, int small_limit, int mid_limit, uint32_t* scan_ticket, uint32_t* scan_total) {
  const uint32_t t = blockIdx.x * kSetupThreads + threadIdx.x;
  const int lane = threadIdx.x & 31;
  uint32_t count = 0;
  CamTri pc;   // the one piece of a triangle the near plane does not cut: registers only
  pc.rt.sign = 0; pc.rt.imin = 0; pc.rt.imax = -1; pc.rt.jmin = 0; pc.rt.jmax = -1; pc.rt.area = 0; pc.rt.mshift = kMacroShiftMin;
#pragma unroll
  for (int k = 0; k < 3; k++) { pc.rt.X[k] = pc.rt.Y[k] = 0; pc.zw[k] = 0.0f; }
  bool mid = false, small = false;
  if (t < n_tris) {
    const DrawRec* dp = draws + find_draw(t, draws, n_draws);
    ClipVert in[3];
    float dn[3];
    const int n_out = cam_vertex_stage(verts, indices, *dp, t, pv.m, in, dn);
    if (n_out == 0) {
      // multi-GPU: a piece whose bounding box touches none of this rank's screen tiles is somebody else's work (a sub-tile triangle of a
      // 4 M-triangle scene touches one or two tiles: seven eighths of them end here on each of eight ranks)
      const bool have = cam_make_piece(in[0], in[1], in[2], W, H, dp->material, pc) &&
                        box_touches_owned_tile<SPLIT>(pc.rt.imin, pc.rt.imax, pc.rt.jmin, pc.rt.jmax, W, tile_rank, tile_nranks);
      uint32_t bs = 0u;
      if (have) {
        const int area = (pc.rt.imax - pc.rt.imin + 1) * (pc.rt.jmax - pc.rt.jmin + 1);
        const bool big = area > mid_limit;
        mid = !big && area > small_limit;
        if (big) {
          const uint32_t slot = atomicAdd(rec_count, 1u);
          if (slot < rec_capacity) {
            if (mid_limit > 0) raster_choose_macro(pc.rt, kMaxItemsManyTris);
            pc.pad = raster_item_count(pc.rt);
            count = pc.pad;
            recs[slot] = pc;
            bs = slot + 1u;
          } else {   // record array full: the warp takes the piece (slow for a wall-sized one, still exact)
            mid = true;
          }
        } else if (!mid) {
          small = true;
        }
      }
      big_slot[t] = bs;
    } else if (n_out < 3) {
      count = cam_setup_clipped<SPLIT>(verts, indices, dp, t, pv.m, W, H, recs, rec_capacity, rec_count, big_slot, vis, tile_rank, tile_nranks, mid_limit);
    } else {
      big_slot[t] = 0u;
    }
  }
  // The block's item scan comes BEFORE the in-line rasterisation: its barriers are then reached by warps that have all done the same
  // (uniform) set-up work, and the divergent pixel loops below end without anybody waiting for the block's slowest warp.
  block_scan_items(count, t, n_tris, item_local, item_block, scan_ticket, scan_total);
  if (small) raster_piece_inline<SPLIT>(pc, t, W, vis, tile_rank, tile_nranks);
  // the mid-sized pieces of the warp's 32 triangles, one after the other, all lanes on each
  for (uint32_t m = __ballot_sync(0xffffffffu, mid); m; m &= m - 1u) raster_piece_warp<SPLIT>(pc, t, __ffs((int)m) - 1, lane, W, vis, tile_rank, tile_nranks);
}

__global__ void __launch_bounds__(256)
cam_raster_kernel(const CamTri* __restrict__ recs, const uint32_t* __restrict__ big_slot, uint32_t n_tris, const uint32_t* __restrict__ item_local,
                  const uint32_t* __restrict__ item_block, uint32_t n_blocks, int W, unsigned long long* __restrict__ vis,
                  const uint32_t* __restrict__ counters, int tile_rank, int tile_nranks) {
  const uint32_t total = counters[CNT_CAM_ITEMS];
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t g = warp; g < total; g += n_warps) {  // one warp per macro tile (8x8 .. 64x64 pixels, per triangle)
    uint32_t rank;
    const uint32_t ti = find_item_triangle(g, item_block, n_blocks, item_local, n_tris, rank);
    uint32_t slot = (__ldg(big_slot + ti) & 0x7FFFFFFFu) - 1u;
    const uint32_t n0 = recs[slot].pad;
    if (rank >= n0) { rank -= n0; slot++; }   // the second piece of a clipped triangle
    const CamTri& v = recs[slot];
    const RasterTri rt = v.rt;
    const float z0 = v.zw[0], z1 = v.zw[1], z2 = v.zw[2];
    MacroItem mi;
    macro_item_setup(rt, rank, lane, mi);
    for (unsigned long long live = mi.live; live; live &= live - 1ull) {
      const int b = __ffsll((long long)live) - 1;
      const int bx0 = mi.x0 + 8 * (b & 7), by0 = mi.y0 + 8 * (b >> 3);
      // multi-GPU: only the 32x32 screen tiles this rank shades need visibility (an 8x8 block lies inside one of them)
      if (!tile_owned(bx0, by0, W, tile_rank, tile_nranks)) continue;
      EdgeBlock eb;
      macro_block_edges(mi, b, eb);
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int p = lane + 32 * h;
        const int i = bx0 + (p & 7), j = by0 + (p >> 3);
        float bc[3];
        if (i >= rt.imin && i <= rt.imax && j >= rt.jmin && j <= rt.jmax && edge_block_sample(eb, p & 7, p >> 3, bc)) {
          const float zw = interp3(bc, z0, z1, z2);
          if (zw >= 0.0f && zw <= 1.0f) {  // near / far (R3); also rejects NaN
            const unsigned long long key = ((unsigned long long)__float_as_uint(zw) << 32) | (unsigned long long)ti;
            atomicMin(&vis[(size_t)j * W + i], key);
          }
        }
      }
    }
  }
}

// the piece of a near-plane-clipped triangle that covers pixel (i, j), for the resolve kernel (rare, not inlined: see cam_clip_pieces)
struct ResolvedPiece { float b[3], iw[3], world[3][3], nn[3][3]; };
__device__ __noinline__ void cam_resolve_clipped(const vct_vertex_t* __restrict__ verts, const uint32_t* __restrict__ indices, const DrawRec* __restrict__ d, uint32_t t,
                                                 const float* __restrict__ pvm, int W, int H, int i, int j, ResolvedPiece* out) {
  CamTri pc[2];
  const int np = cam_clip_pieces(verts, indices, d, t, pvm, W, H, pc);
  float b[3] = {0.f, 0.f, 0.f};
  const int q = (np > 1 && !raster_sample(pc[0].rt, i, j, b)) ? 1 : 0;
  raster_sample(pc[q].rt, i, j, b);
  for (int k = 0; k < 3; k++) {
    out->b[k] = b[k]; out->iw[k] = pc[q].iw[k];
    for (int c = 0; c < 3; c++) { out->world[k][c] = pc[q].world[k][c]; out->nn[k][c] = pc[q].nn[k][c]; }
  }
}

// key of the cleared depth buffer: depth 1.0 (glClear), no triangle.  A fragment at zw == 1.0 has a
// smaller key only if its triangle id is smaller than 0xFFFFFFFF, so mask it explicitly below.
constexpr unsigned long long kVisClear = ((unsigned long long)0x3F800000u << 32) | 0xFFFFFFFFull;

// what the resolve kernel needs to build the cone tracer's live-tile list on the way (optional: list == nullptr)
struct TileListOut {
  uint32_t* list;          // [n] 8x4 tiles (index ty * tiles_x + tx) that hold at least one pixel the cone tracer shades
  uint32_t* count;         // zeroed by the clear kernel of this pass
  uint32_t* work_counter;  // the persistent cone kernel's work counter: reset here for the frame
  float cube_size;
};

// the piece of the winning triangle that covers pixel (i, j) when the triangle has no record (or the record array overflowed): vertex stage
// again; out of line in the LEAN instantiation (small scenes: every triangle has a record, the path is dead weight in registers there)
__device__ __forceinline__ void cam_resolve_recompute(const vct_vertex_t* __restrict__ verts, const uint32_t* __restrict__ indices, const DrawRec* __restrict__ dp, uint32_t ti,
                                                      const float* __restrict__ pvm, int W, int H, int i, int j, float (&b)[3], float (&iw)[3], float (&world)[3][3],
                                                      float (&nn)[3][3]) {
  ClipVert in[3];
  float dn[3];
  const int n_out = cam_vertex_stage(verts, indices, *dp, ti, pvm, in, dn);
  if (n_out == 0) {   // registers only
    CamTri pc;
    cam_make_piece(in[0], in[1], in[2], W, H, dp->material, pc);
    raster_sample(pc.rt, i, j, b);
#pragma unroll
    for (int k = 0; k < 3; k++) {
      iw[k] = pc.iw[k];
#pragma unroll
      for (int c = 0; c < 3; c++) { world[k][c] = pc.world[k][c]; nn[k][c] = pc.nn[k][c]; }
    }
  } else {
    ResolvedPiece rp;
    cam_resolve_clipped(verts, indices, dp, ti, pvm, W, H, i, j, &rp);
#pragma unroll
    for (int k = 0; k < 3; k++) {
      b[k] = rp.b[k]; iw[k] = rp.iw[k];
#pragma unroll
      for (int c = 0; c < 3; c++) { world[k][c] = rp.world[k][c]; nn[k][c] = rp.nn[k][c]; }
    }
  }
}
__device__ __noinline__ void cam_resolve_recompute_out_of_line(const vct_vertex_t* __restrict__ verts, const uint32_t* __restrict__ indices, const DrawRec* __restrict__ dp,
                                                               uint32_t ti, const float* __restrict__ pvm, int W, int H, int i, int j, ResolvedPiece* out) {
  float b[3], iw[3], world[3][3], nn[3][3];
  cam_resolve_recompute(verts, indices, dp, ti, pvm, W, H, i, j, b, iw, world, nn);
  for (int k = 0; k < 3; k++) {
    out->b[k] = b[k]; out->iw[k] = iw[k];
    for (int c = 0; c < 3; c++) { out->world[k][c] = world[k][c]; out->nn[k][c] = nn[k][c]; }
  }
}

// One warp per 8 x 4 pixel tile (the cone tracer's unit), kResolveTiles consecutive tiles per warp: world position and interpolated normal
// of the winning triangle per pixel, and -- fused here, it was a launch of its own -- the list of tiles with at least one pixel to shade
// (voxel_cone_tracing.frag:248-251: hit, and inside the grid cube), appended with one atomic per warp.
constexpr int kResolveTiles = 4;
template <bool LEAN>
__global__ void __launch_bounds__(256, LEAN ? 4 : 1)   // LEAN: 64 registers -- the out-of-line recompute path (dead for small scenes) may spill
cam_resolve_kernel(const vct_vertex_t* __restrict__ verts, const uint32_t* __restrict__ indices, const DrawRec* __restrict__ draws, uint32_t n_draws, Mat4 pv,
                   const CamTri* __restrict__ recs, const uint32_t* __restrict__ big_slot, const unsigned long long* vis, int W, int H, float* __restrict__ world_pos,
                   float* __restrict__ normal, uint32_t* __restrict__ material, unsigned long long* vis_out, int tile_rank, int tile_nranks, const TileListOut tl) {
  const int lane = threadIdx.x & 31;
  if (tl.list && blockIdx.x == 0 && threadIdx.x == 0) *tl.work_counter = 0u;   // the cone kernel of this frame starts at item 0
  const int tiles_x = (W + 7) / 8, tiles_y = (H + 3) / 4, n_tiles = tiles_x * tiles_y;
  const int n_groups = (n_tiles + kResolveTiles - 1) / kResolveTiles;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int grp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; grp < n_groups; grp += warps) {
    uint32_t live_mask = 0;
#pragma unroll 1
    for (int k = 0; k < kResolveTiles; k++) {
      const int tile = grp * kResolveTiles + k;
      if (tile >= n_tiles) break;
      const int tile_x = tile % tiles_x, tile_y = tile / tiles_x;
      const int i = tile_x * 8 + (lane & 7), j = tile_y * 4 + (lane >> 3);
      bool live = false;
      if (i < W && j < H && tile_owned(i, j, W, tile_rank, tile_nranks)) {   // pixels of other ranks' tiles are never read by this rank's tracer
        const size_t px = (size_t)j * W + i;
        const unsigned long long key = vis[px];
        uint32_t ti = (uint32_t)(key & 0xFFFFFFFFull);
        // GL_LESS against the cleared depth 1.0: a fragment exactly at zw == 1.0 fails
        if ((uint32_t)(key >> 32) >= 0x3F800000u) ti = VCT_NO_TRIANGLE;
        if (ti == VCT_NO_TRIANGLE) {
          if (vis_out) vis_out[px] = kVisClear;
          material[px] = VCT_NO_TRIANGLE;
        } else {
          // the piece of the winning triangle that covers this pixel: from its records, or by running the vertex stage again
          float b[3], iw[3], world[3][3], nn[3][3];
          uint32_t mat = 0;
          bool found = false;
          const uint32_t bs = __ldg(big_slot + ti);
          if (bs) {
            uint32_t slot = (bs & 0x7FFFFFFFu) - 1u;
            found = raster_sample(recs[slot].rt, i, j, b);
            if (!found && (bs >> 31)) { slot++; found = raster_sample(recs[slot].rt, i, j, b); }
            if (found) {
              const CamTri& v = recs[slot];
#pragma unroll
              for (int q = 0; q < 3; q++) {
                iw[q] = v.iw[q];
#pragma unroll
                for (int c = 0; c < 3; c++) { world[q][c] = v.world[q][c]; nn[q][c] = v.nn[q][c]; }
              }
              mat = v.material;
            }
          }
          if (!found) {   // no record (sub-tile triangle), or the pixel belongs to a piece that was rasterised in line
            const DrawRec* dp = draws + find_draw(ti, draws, n_draws);
            if (LEAN) {
              ResolvedPiece rp;
              cam_resolve_recompute_out_of_line(verts, indices, dp, ti, pv.m, W, H, i, j, &rp);
#pragma unroll
              for (int q = 0; q < 3; q++) {
                b[q] = rp.b[q]; iw[q] = rp.iw[q];
#pragma unroll
                for (int c = 0; c < 3; c++) { world[q][c] = rp.world[q][c]; nn[q][c] = rp.nn[q][c]; }
              }
            } else {
              cam_resolve_recompute(verts, indices, dp, ti, pv.m, W, H, i, j, b, iw, world, nn);
            }
            mat = dp->material;
          }
          const float q3[3] = {b[0] * iw[0], b[1] * iw[1], b[2] * iw[2]};
          const float qs = (q3[0] + q3[1]) + q3[2];
          float wp[3];
#pragma unroll
          for (int c = 0; c < 3; c++) {
            wp[c] = interp3(q3, world[0][c], world[1][c], world[2][c]) / qs;
            world_pos[px * 3 + c] = wp[c];
            normal[px * 3 + c] = interp3(q3, nn[0][c], nn[1][c], nn[2][c]) / qs;
          }
          material[px] = mat;
          // the tracer's within_cube test (load_pixel in cone_trace.cu: same expression, same inputs)
          const float p0 = 0.5f * (wp[0] / tl.cube_size) + 0.5f, p1 = 0.5f * (wp[1] / tl.cube_size) + 0.5f, p2 = 0.5f * (wp[2] / tl.cube_size) + 0.5f;
          live = fabsf(p0) < 1.0f && fabsf(p1) < 1.0f && fabsf(p2) < 1.0f;
        }
      }
      if (__any_sync(0xffffffffu, live)) live_mask |= 1u << k;
    }
    if (tl.list && live_mask) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(tl.count, (uint32_t)__popc(live_mask));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (lane < kResolveTiles && ((live_mask >> lane) & 1u)) tl.list[base + __popc(live_mask & ((1u << lane) - 1u))] = (uint32_t)(grp * kResolveTiles + lane);
    }
  }
}

__global__ void fill_u64_kernel(unsigned long long* p, size_t n, unsigned long long v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
// the camera pass's clear: the visibility buffer plus its two counters (record count, live-tile count) in one launch
__global__ void cam_clear_kernel(unsigned long long* p, size_t n, unsigned long long v, uint32_t* c0, uint32_t* c1) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
  if (blockIdx.x == 0 && threadIdx.x == 0) { if (c0) *c0 = 0u; if (c1) *c1 = 0u; }
}
__global__ void fill_u32_kernel(uint32_t* p, size_t n, uint32_t v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void copy_u32x4_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n4, uint32_t* dst_tail, const uint32_t* src_tail, int n_tail) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
  if (blockIdx.x == 0 && (int)threadIdx.x < n_tail) dst_tail[threadIdx.x] = src_tail[threadIdx.x];
}
// device-to-device copy of n words by a kernel (16-byte accesses): a cudaMemcpyAsync would go through a copy engine
int launch_copy_u32(cudaStream_t s, uint32_t* dst, const uint32_t* src, size_t n) {
  if (n == 0) return VCT_OK;
  const size_t n4 = n / 4;
  copy_u32x4_kernel<<<grid_for(n4 ? n4 : 1, 256, 148 * 8), 256, 0, s>>>((uint4*)dst, (const uint4*)src, n4, dst + 4 * n4, src + 4 * n4, (int)(n - 4 * n4));
  VCT_CUDA(cudaGetLastError());
  return VCT_OK;
}
int launch_fill_u32(cudaStream_t s, uint32_t* p, size_t n, uint32_t v) {
  if (n == 0) return VCT_OK;
  fill_u32_kernel<<<grid_for(n), 256, 0, s>>>(p, n, v);
  VCT_CUDA(cudaGetLastError());
  return VCT_OK;
}

static void mat4_mul_host(const float* a, const float* b, float* out) {  // column-major a*b, no contraction on the host
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 4; r++) {
      volatile float s0 = a[0 * 4 + r] * b[c * 4 + 0];
      volatile float s1 = a[1 * 4 + r] * b[c * 4 + 1];
      volatile float s2 = a[2 * 4 + r] * b[c * 4 + 2];
      volatile float s3 = a[3 * 4 + r] * b[c * 4 + 3];
      volatile float t = s0 + s1;
      t = t + s2;
      out[c * 4 + r] = t + s3;
    }
}

int launch_gbuffer(vct_device* dev, vct_scene* sc, const float* view, const float* proj, vct_target_t_* t, int tile_rank, int tile_nranks, bool tile_list) {
  cudaStream_t s = dev->stream;
  if (tile_nranks < 1) { tile_nranks = 1; tile_rank = 0; }
  tile_rank |= screen_tile_k(tile_nranks) << 16;   // the kernels' tile_owned() unpacks it
  const size_t npx = (size_t)t->W * t->H;
  const int n_tiles = ((t->W + 7) / 8) * ((t->H + 3) / 4);
  if (tile_list && !t->tile_list) VCT_CUDA(cudaMalloc(&t->tile_list, ((size_t)n_tiles + 1) * sizeof(uint32_t)));
  TileListOut tl;
  memset(&tl, 0, sizeof tl);
  tl.cube_size = sc->cube_size;
  if (tile_list) { tl.list = t->tile_list + 1; tl.count = t->tile_list; tl.work_counter = dev->counters + CNT_CONE_WORK; }
  uint32_t* rec_count = dev->counters + CNT_CAM_RECS;
  cam_clear_kernel<<<dev->prop.multiProcessorCount * 8, 256, 0, s>>>(t->vis, npx, kVisClear, rec_count, tl.count);
  const int sms = dev->prop.multiProcessorCount;
  if (sc->n_tris) {
    // records only for the pieces that become work items: a compact array (a quarter of the triangles, at least 64 k); a scene
    // that fills it falls back to the in-line path for the rest
    const size_t rec_capacity = sc->n_tris / 4 > 65536 ? sc->n_tris / 4 : 65536;
    int rc = ensure_tri_scratch(dev, 1, sc->n_tris, rec_capacity * sizeof(CamTri));
    if (rc) return rc;
    Mat4 pv;
    mat4_mul_host(proj, view, pv.m);  // projection * view (voxel_cone_tracing.vert:25)
    const uint32_t n_blocks = (sc->n_tris + kSetupThreads - 1) / kSetupThreads;
    CamTri* recs = (CamTri*)dev->rs[1].tri_recs;
    const bool many = sc->n_tris >= kSmallPathMinTris;
    if (tile_nranks > 1)
      cam_setup_kernel<true><<<n_blocks, kSetupThreads, 0, s>>>(sc->verts, sc->indices, sc->draws, sc->n_draws, sc->n_tris, pv, t->W, t->H, recs, (uint32_t)rec_capacity, rec_count,
                                                        dev->rs[1].big_slot, dev->rs[1].item_local, dev->rs[1].item_block, t->vis, tile_rank, tile_nranks,
                                                        many ? (dev->debug_small_limit >= 0 ? dev->debug_small_limit : kSmallCamPixels) : 0, many ? kMidCamPixels : 0, dev->counters + CNT_TICKET_CAM,
                                                        dev->counters + CNT_CAM_ITEMS);
    else
      cam_setup_kernel<false><<<n_blocks, kSetupThreads, 0, s>>>(sc->verts, sc->indices, sc->draws, sc->n_draws, sc->n_tris, pv, t->W, t->H, recs, (uint32_t)rec_capacity, rec_count,
                                                        dev->rs[1].big_slot, dev->rs[1].item_local, dev->rs[1].item_block, t->vis, tile_rank, tile_nranks,
                                                        many ? (dev->debug_small_limit >= 0 ? dev->debug_small_limit : kSmallCamPixels) : 0, many ? kMidCamPixels : 0, dev->counters + CNT_TICKET_CAM,
                                                        dev->counters + CNT_CAM_ITEMS);
    cam_raster_kernel<<<sms * 8, 256, 0, s>>>(recs, dev->rs[1].big_slot, sc->n_tris, dev->rs[1].item_local, dev->rs[1].item_block, n_blocks, t->W, t->vis, dev->counters,
                                              tile_rank, tile_nranks);
    const int groups = (n_tiles + kResolveTiles - 1) / kResolveTiles;
    const int blocks = min((groups + 7) / 8, sms * 8);
    if (many)
      cam_resolve_kernel<false><<<blocks, 256, 0, s>>>(sc->verts, sc->indices, sc->draws, sc->n_draws, pv, recs, dev->rs[1].big_slot, t->vis, t->W, t->H, t->world_pos,
                                                       t->normal, t->material, t->vis, tile_rank, tile_nranks, tl);
    else
      cam_resolve_kernel<true><<<blocks, 256, 0, s>>>(sc->verts, sc->indices, sc->draws, sc->n_draws, pv, recs, dev->rs[1].big_slot, t->vis, t->W, t->H, t->world_pos,
                                                      t->normal, t->material, t->vis, tile_rank, tile_nranks, tl);
  } else {
    launch_fill_u32(s, t->material, npx, VCT_NO_TRIANGLE);
    if (tile_list) launch_fill_u32(s, dev->counters + CNT_CONE_WORK, 1, 0u);
  }
  VCT_CUDA(cudaGetLastError());
  return VCT_OK;
}

}  // namespace vct
