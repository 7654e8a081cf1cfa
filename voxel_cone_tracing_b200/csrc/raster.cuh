// raster.cuh -- triangle setup, coverage and work-item machinery shared by the voxelizer and
// the camera (G-buffer) pass.  Implements raster rules R1-R3 of DESIGN.md: vertices snapped to
// 1/256 pixel, 64-bit integer edge functions at pixel centres, top-left rule, barycentrics
// b_k = float(E_k) / float(2A).  Replaces the fixed-function rasteriser the reference drives at
// src/renderer.cpp:339-347 (voxelization viewport 2R x 2R) and :358-389 (camera viewport).
//
// Work distribution: every triangle's bounding box is cut into square macro tiles ("items") of 2^s pixels,
// s = 3..6 chosen PER TRIANGLE as the smallest size that gives it at most kMaxItemsPerTri items; a block-local
// scan in the setup kernel plus a one-block scan of the block totals give every item a global index
// without a host round trip; the raster kernel walks items with one warp per item: the warp first
// rejects the 8x8-pixel blocks of the macro tile that no edge function can reach (two blocks per lane,
// one ballot), then visits the live blocks with 2 pixels per lane.
// Why per triangle: with 8x8 items only, a wall at 8K is half a million items, each paying two binary
// searches through the prefix arrays (17 of the 33 ms of the config-5 G-buffer pass in round 1); with
// 64x64 items only, the 1112 triangles of the Cornell box are ~200 items of up to 64 serial blocks and
// nine tenths of the resident warps idle (vox_raster_kernel 25 -> 298 us at 256^3).
#pragma once

#include "vct_internal.cuh"

namespace vct {

constexpr int kSetupThreads = 256;
constexpr int kTile = 8;  // 8x8 pixels per block: 2 pixels per lane
constexpr int kMacroShiftMin = 3, kMacroShiftMax = 6;  // a work item is 8x8 .. 64x64 pixels = 1 .. 64 blocks
constexpr uint32_t kMaxItemsPerTri = 1024;
// The in-thread small-triangle path of the setup kernels pays off when there are many triangles; a scene of a few thousand
// triangles has too few setup threads to hide the fragment work there, its 8x8 items parallelise better.
constexpr uint32_t kSmallPathMinTris = 32768;
__device__ __forceinline__ int imin3(int a, int b, int c) { return min(a, min(b, c)); }
__device__ __forceinline__ int imax3(int a, int b, int c) { return max(a, max(b, c)); }

// R1/R2. Returns false (t.sign = 0) when the triangle produces no fragment.
__device__ __forceinline__ bool raster_setup(const float xw[3], const float yw[3], int W, int H, RasterTri& t) {
  t.sign = 0;
  t.imin = 0; t.imax = -1; t.jmin = 0; t.jmax = -1;
  t.area = 0;
  t.mshift = kMacroShiftMin;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    if (!(fabsf(xw[k]) <= 2097152.0f) || !(fabsf(yw[k]) <= 2097152.0f)) return false;  // guard band / NaN
    t.X[k] = (int)rintf(xw[k] * 256.0f);
    t.Y[k] = (int)rintf(yw[k] * 256.0f);
  }
  long long a = ((long long)t.X[1] - t.X[0]) * ((long long)t.Y[2] - t.Y[0]) -
                ((long long)t.Y[1] - t.Y[0]) * ((long long)t.X[2] - t.X[0]);
  if (a == 0) return false;
  int minx = imin3(t.X[0], t.X[1], t.X[2]), maxx = imax3(t.X[0], t.X[1], t.X[2]);
  int miny = imin3(t.Y[0], t.Y[1], t.Y[2]), maxy = imax3(t.Y[0], t.Y[1], t.Y[2]);
  // pixel centres (i*256+128) inside [min,max]:  i >= ceil((min-128)/256), i <= floor((max-128)/256)
  int i0 = (minx - 128 + 255) >> 8, i1 = (maxx - 128) >> 8;
  int j0 = (miny - 128 + 255) >> 8, j1 = (maxy - 128) >> 8;
  i0 = max(i0, 0); j0 = max(j0, 0);
  i1 = min(i1, W - 1); j1 = min(j1, H - 1);
  if (i0 > i1 || j0 > j1) return false;
  t.imin = i0; t.imax = i1; t.jmin = j0; t.jmax = j1;
  t.area = a > 0 ? a : -a;
  t.sign = a > 0 ? 1 : -1;
  int ms = kMacroShiftMin;
  while (ms < kMacroShiftMax && (uint32_t)((i1 >> ms) - (i0 >> ms) + 1) * (uint32_t)((j1 >> ms) - (j0 >> ms) + 1) > kMaxItemsPerTri) ms++;
  t.mshift = ms;
  return true;
}

// Scenes of many triangles (the set-up kernels' in-line paths are on: mid_limit > 0) want FEW, large work items per triangle: an item costs
// two binary searches, a 160-byte record load and the block culling before its first pixel, and the 4 M-triangle scene at 8K turned its
// triangles of 70..200 pixels across into 977 k items of 8 x 8 pixels (cam_raster_kernel 2.4 ms, 1.5 G warp instructions).  Small scenes
// keep up to kMaxItemsPerTri small items per triangle: their few triangles need the parallelism.
constexpr uint32_t kMaxItemsManyTris = 16;
__device__ __forceinline__ void raster_choose_macro(RasterTri& t, uint32_t max_items) {
  int ms = kMacroShiftMin;
  while (ms < kMacroShiftMax && (uint32_t)((t.imax >> ms) - (t.imin >> ms) + 1) * (uint32_t)((t.jmax >> ms) - (t.jmin >> ms) + 1) > max_items) ms++;
  t.mshift = ms;
}

__device__ __forceinline__ uint32_t raster_item_count(const RasterTri& t) {
  if (t.sign == 0) return 0u;
  const int ms = t.mshift;
  int tx = (t.imax >> ms) - (t.imin >> ms) + 1, ty = (t.jmax >> ms) - (t.jmin >> ms) + 1;
  return (uint32_t)tx * (uint32_t)ty;
}

// R2/R3: coverage + barycentrics of pixel (i,j)
__device__ __forceinline__ bool raster_sample(const RasterTri& t, int i, int j, float b[3]) {
  long long px = (long long)i * 256 + 128, py = (long long)j * 256 + 128;
  long long E[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const int a = (k + 1) % 3, c = (k + 2) % 3;
    long long dx = (long long)t.X[c] - t.X[a], dy = (long long)t.Y[c] - t.Y[a];
    long long e = dx * (py - t.Y[a]) - dy * (px - t.X[a]);
    if (t.sign < 0) { e = -e; dx = -dx; dy = -dy; }
    if (e < 0) return false;
    if (e == 0 && !((dy < 0) || (dy == 0 && dx < 0))) return false;  // top-left rule
    E[k] = e;
  }
  float fa = (float)t.area;
  b[0] = (float)E[0] / fa;
  b[1] = (float)E[1] / fa;
  b[2] = (float)E[2] / fa;
  return true;
}

// The same coverage test and barycentrics as raster_sample, factored for a block of pixels: the three edge functions are
// evaluated once (64-bit) at pixel (i0,j0); a pixel at offset (ox,oy) then costs two 32x32->64 multiply-adds per edge.
// E_k(px + 256*ox, py + 256*oy) = E_k(px,py) + dx_k*256*oy - dy_k*256*ox -- exact integers, identical to raster_sample.
struct EdgeBlock {
  long long e0[3];
  int dx[3], dy[3];   // after the orientation flip
  float fa;
};
__device__ __forceinline__ void edge_block_setup(const RasterTri& t, int i0, int j0, EdgeBlock& q) {
  const long long px = (long long)i0 * 256 + 128, py = (long long)j0 * 256 + 128;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const int a = (k + 1) % 3, c = (k + 2) % 3;
    int dx = t.X[c] - t.X[a], dy = t.Y[c] - t.Y[a];   // |X|,|Y| <= 2^29: no overflow
    long long e = (long long)dx * (py - t.Y[a]) - (long long)dy * (px - t.X[a]);
    if (t.sign < 0) { e = -e; dx = -dx; dy = -dy; }
    q.e0[k] = e; q.dx[k] = dx; q.dy[k] = dy;
  }
  q.fa = (float)t.area;
}
__device__ __forceinline__ bool edge_block_sample(const EdgeBlock& q, int ox, int oy, float b[3]) {
  long long E[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const long long e = q.e0[k] + (long long)q.dx[k] * (256 * oy) - (long long)q.dy[k] * (256 * ox);
    if (e < 0) return false;
    if (e == 0 && !((q.dy[k] < 0) || (q.dy[k] == 0 && q.dx[k] < 0))) return false;  // top-left rule
    E[k] = e;
  }
  b[0] = (float)E[0] / q.fa;
  b[1] = (float)E[1] / q.fa;
  b[2] = (float)E[2] / q.fa;
  return true;
}

// ---- one work item = macro tile `rank` of the triangle's bounding box ----
struct MacroItem {
  int x0, y0;              // pixel origin of the macro tile
  EdgeBlock em;            // edge functions at that origin
  unsigned long long live; // bit (by * 8 + bx): block (bx, by) may hold covered pixel centres
};
// all 32 lanes call it
__device__ __forceinline__ void macro_item_setup(const RasterTri& rt, uint32_t rank, int lane, MacroItem& m) {
  const int ms = rt.mshift, nb = 1 << (ms - 3);   // the macro tile is nb x nb blocks; block index b = by * 8 + bx whatever nb is
  const int tiles_x = (rt.imax >> ms) - (rt.imin >> ms) + 1;
  const int tx = (rt.imin >> ms) + (int)(rank % (uint32_t)tiles_x), ty = (rt.jmin >> ms) + (int)(rank / (uint32_t)tiles_x);
  m.x0 = tx << ms; m.y0 = ty << ms;
  edge_block_setup(rt, m.x0, m.y0, m.em);
  uint32_t mask[2];
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int b = lane + 32 * h, bx = b & 7, by = b >> 3;
    const int ox = m.x0 + 8 * bx, oy = m.y0 + 8 * by;
    bool ok = bx < nb && by < nb && ox <= rt.imax && ox + 7 >= rt.imin && oy <= rt.jmax && oy + 7 >= rt.jmin;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      // the largest value edge k takes on the pixel centres of the block (corners): negative = the block is outside
      const long long e = m.em.e0[k] + (long long)m.em.dx[k] * (256 * 8 * by) - (long long)m.em.dy[k] * (256 * 8 * bx);
      const long long emax = e + (m.em.dx[k] > 0 ? (long long)m.em.dx[k] * (256 * 7) : 0ll) + (m.em.dy[k] < 0 ? -(long long)m.em.dy[k] * (256 * 7) : 0ll);
      ok = ok && emax >= 0;
    }
    mask[h] = __ballot_sync(0xffffffffu, ok);
  }
  m.live = ((unsigned long long)mask[1] << 32) | mask[0];
}
// edge functions at the origin of block b of the macro tile
__device__ __forceinline__ void macro_block_edges(const MacroItem& m, int b, EdgeBlock& eb) {
  const int bx = b & 7, by = b >> 3;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    eb.e0[k] = m.em.e0[k] + (long long)m.em.dx[k] * (256 * 8 * by) - (long long)m.em.dy[k] * (256 * 8 * bx);
    eb.dx[k] = m.em.dx[k]; eb.dy[k] = m.em.dy[k];
  }
  eb.fa = m.em.fa;
}

__device__ __forceinline__ float interp3(const float b[3], float a0, float a1, float a2) {
  return (b[0] * a0 + b[1] * a1) + b[2] * a2;  // compiled with -fmad=false: two roundings per term, like the oracle
}

// ---- exclusive scan of per-triangle item counts, inside the setup kernels ----
// Every block writes local[t] (exclusive prefix inside the 256-thread block) and block_total[blockIdx.x]; the LAST block to
// finish (ticket counter, reset for the next launch) turns the block totals into exclusive prefixes in place and stores the
// grand total -- the work of a separate one-block kernel (4.5 us of launch + tail per rasteriser) folded into the setup kernel.
__device__ __forceinline__ void block_scan_items(uint32_t count, uint32_t t, uint32_t n_tris, uint32_t* local, uint32_t* block_total, uint32_t* ticket,
                                                 uint32_t* total) {
  __shared__ uint32_t warp_sums[kSetupThreads / 32];
  __shared__ uint32_t carry_s, is_last;
  constexpr int kWarps = kSetupThreads / 32;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t v = count;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  if (lane == 31) warp_sums[wid] = v;
  __syncthreads();
  if (wid == 0) {
    uint32_t w = lane < kWarps ? warp_sums[lane] : 0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t n = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += n;
    }
    if (lane < kWarps) warp_sums[lane] = w;  // inclusive
  }
  __syncthreads();
  uint32_t warp_off = wid ? warp_sums[wid - 1] : 0u;
  if (t < n_tris) local[t] = warp_off + v - count;
  if (threadIdx.x == kSetupThreads - 1) {
    block_total[blockIdx.x] = warp_off + v;
    __threadfence();                                       // the total is visible before the ticket is taken
    is_last = atomicAdd(ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
    carry_s = 0;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const uint32_t n_blocks = gridDim.x;
  for (uint32_t base = 0; base < n_blocks; base += kSetupThreads) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t c = i < n_blocks ? __ldcg(block_total + i) : 0u;   // written by other blocks: read from L2
    uint32_t x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t n = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += n;
    }
    __syncthreads();                                       // warp_sums of the previous round / of the block scan are consumed
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
      uint32_t w = lane < kWarps ? warp_sums[lane] : 0u;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += n;
      }
      if (lane < kWarps) warp_sums[lane] = w;
    }
    __syncthreads();
    const uint32_t off = carry_s + (wid ? warp_sums[wid - 1] : 0u);
    if (i < n_blocks) block_total[i] = off + x - c;
    __syncthreads();
    if (threadIdx.x == kSetupThreads - 1) carry_s = off + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) { *total = carry_s; *ticket = 0u; }
}

// item index -> (triangle, tile).  item_block: exclusive prefix per setup block; item_local: exclusive
// prefix inside the block.  Picks the LAST entry whose prefix <= the target so that empty triangles /
// empty blocks (equal prefixes) are skipped.
__device__ __forceinline__ uint32_t find_item_triangle(uint32_t g, const uint32_t* __restrict__ item_block, uint32_t n_blocks,
                                                       const uint32_t* __restrict__ item_local, uint32_t n_tris, uint32_t& rank) {
  uint32_t lo = 0, hi = n_blocks;  // invariant: item_block[lo] <= g
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (__ldg(item_block + mid) <= g) lo = mid; else hi = mid;
  }
  uint32_t x = g - __ldg(item_block + lo);
  uint32_t tlo = lo * kSetupThreads, thi = min(tlo + kSetupThreads, n_tris);
  while (thi - tlo > 1) {
    uint32_t mid = (tlo + thi) >> 1;
    if (__ldg(item_local + mid) <= x) tlo = mid; else thi = mid;
  }
  rank = x - __ldg(item_local + tlo);
  return tlo;
}

// draw lookup for a global triangle sequence number
__device__ __forceinline__ uint32_t find_draw(uint32_t tri, const DrawRec* __restrict__ draws, uint32_t n_draws) {
  uint32_t lo = 0, hi = n_draws;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (draws[mid].tri_base <= tri) lo = mid; else hi = mid;
  }
  return lo;
}

}  // namespace vct
