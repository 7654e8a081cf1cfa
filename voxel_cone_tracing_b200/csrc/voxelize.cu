// voxelize.cu -- triangle voxelization into the RGBA8 radiance/opacity grid (level 0).
//
// Replaces the reference's voxelize.vert / voxelize.geom / fixed-function raster / voxelize.frag
// pass (src/renderer.cpp:316-347).  Not a port: the geometry-shader + ROP pipeline becomes
//   1. vox_setup_kernel   triangle-parallel: vertex transform (voxelize.vert:24-30), dominant-axis
//                         selection (voxelize.geom:25-55), 1/256-pixel snapping, 8x8-pixel item count,
//                         block-local scan
//   2. (inside 1.)        the last block to finish scans the block totals: global item offsets (no host round trip, no extra launch)
//   3. vox_raster_kernel  one warp per macro tile (8x8 .. 64x64 pixels, per triangle) (8x8 blocks culled by the edge functions): coverage, fragment shading (voxelize.frag:122-157),
//                         append of a 32-byte fragment record to a per-voxel linked list whose head
//                         lives in the grid word itself (atomicExch); the fragment that finds the voxel
//                         empty marks its arena slot (`fresh`).  Triangles of <= 100 pixels are rasterised
//                         inside 1. (count pass, one arena atomic per warp, write pass).
//   4. vox_resolve_kernel one thread per arena slot, the `fresh` ones resolve their voxel: sorts the voxel's fragments by the canonical
//                         order key (draw, triangle, row, column) and folds them with the reference's
//                         RGBA8 running average (voxelize.frag:95-120) -> deterministic and bit-exact
//                         against the sequential oracle, which the CAS loop of the reference is not.
//                         Also marks the 32x8x8 tile of the voxel (sparse mip build) and, multi-GPU, stores
//                         the voxel into every peer's grid over NVLink.
//   sparse_clear_kernel   vct_grid_clear in a frame loop: zeroes the voxels of the previous frame's
//                         `fresh` fragments instead of the whole level.
// Built with -fmad=false: the arithmetic (IEEE add/mul/div/sqrt only, fixed evaluation order) is the
// same as the oracle's so that voxel occupancy AND colour match bit for bit.
#include <cuda_fp16.h>

#include "raster.cuh"

namespace vct {

struct F3 { float x, y, z; };
__device__ __forceinline__ F3 f3(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ F3 sub(F3 a, F3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float dot(F3 a, F3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ F3 cross(F3 a, F3 b) { return f3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
__device__ __forceinline__ F3 normalize(F3 a) { float l = sqrtf(dot(a, a)); return f3(a.x / l, a.y / l, a.z / l); }
__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

// fragment colour (voxelize.frag:122-153) at interpolated position `pos`, returns colour * 255
__device__ __forceinline__ void shade_fragment(const VoxTri& v, const float b[3], const vct_material_t* __restrict__ mats, const Lights& L,
                                               float cube_size, F3 pos, float val[4]) {
  F3 nrm = f3(interp3(b, v.nn[0][0], v.nn[1][0], v.nn[2][0]), interp3(b, v.nn[0][1], v.nn[1][1], v.nn[2][1]),
              interp3(b, v.nn[0][2], v.nn[1][2], v.nn[2][2]));
  F3 color = f3(0.f, 0.f, 0.f);
  for (int i = 0; i < L.n; i++) {
    F3 lp = f3(L.l[i].position[0] / cube_size, L.l[i].position[1] / cube_size, L.l[i].position[2] / cube_size);
    F3 dv = sub(lp, pos);
    float dist = sqrtf(dot(dv, dv));
    F3 dir = f3(dv.x / dist, dv.y / dist, dv.z / dist);
    float att = 1.0f / ((1.0f + 0.0f * dist) + (1.0f * dist) * dist);
    float cos_surf = fmaxf(dot(normalize(nrm), dir), 0.0f);
    float s = cos_surf * att;
    color.x = color.x + (L.l[i].color[0] * s) * L.l[i].intensity;
    color.y = color.y + (L.l[i].color[1] * s) * L.l[i].intensity;
    color.z = color.z + (L.l[i].color[2] * s) * L.l[i].intensity;
  }
  const vct_material_t& m = mats[v.material];
  color = f3(m.diffuse[0] * color.x + m.emission[0], m.diffuse[1] * color.y + m.emission[1], m.diffuse[2] * color.z + m.emission[2]);
  float tr0 = 1.f, tr1 = 1.f, tr2 = 1.f, alpha = 1.f;
  if (m.illum == 4 || m.illum == 6 || m.illum == 7 || m.illum == 9) {
    tr0 = m.transmittance[0]; tr1 = m.transmittance[1]; tr2 = m.transmittance[2];
    alpha = m.dissolve;
  }
  val[0] = clamp01(tr0 * color.x) * 255.0f;
  val[1] = clamp01(tr1 * color.y) * 255.0f;
  val[2] = clamp01(tr2 * color.z) * 255.0f;
  val[3] = clamp01(alpha) * 255.0f;
}

// what the fragment stage needs besides the triangle
struct FragCtx {
  const vct_material_t* mats;
  Lights L;
  float cube_size;
  int R, z0, z1;
  uint32_t* base;
  FragRec* frags;
  uint32_t frag_capacity;
  uint8_t* fresh;      // per arena slot: this fragment was the first of its voxel
  uint32_t* counters;
  // VCT_ACCUM_FIXED_POINT (non-reference variant): two 64-bit accumulators per arena slot; the slot of a voxel's FIRST fragment collects
  // the whole voxel: word 0 = sum R | sum G << 24 | count << 48, word 1 = sum B | sum A << 24.  nullptr = the reference's ordered mode.
  unsigned long long* accum;
  int vstride;         // 32-bit words per voxel of `base`: 1 (RGBA8) or 2 (RGBA16F: the low word doubles as the claim word until the resolve pass)
};

// Interpolated position of a covered pixel and its voxel (voxelize.frag:156-157: truncation, then the image bounds check;
// multi-GPU: the z-slab test).  false = the fragment writes nothing.  The ONE place where this is decided: the counting pass and
// the writing pass of the rasterisers must agree exactly.
__device__ __forceinline__ bool fragment_voxel(const FragCtx& c, const VoxTri& v, const float b[3], F3& pos, uint32_t& voxel) {
  pos = f3(interp3(b, v.wp[0][0], v.wp[1][0], v.wp[2][0]), interp3(b, v.wp[0][1], v.wp[1][1], v.wp[2][1]),
           interp3(b, v.wp[0][2], v.wp[1][2], v.wp[2][2]));
  const float fR = (float)c.R;
  const int vx = (int)(fR * (0.5f * pos.x + 0.5f)), vy = (int)(fR * (0.5f * pos.y + 0.5f)), vz = (int)(fR * (0.5f * pos.z + 0.5f));
  voxel = ((uint32_t)vz * (uint32_t)c.R + (uint32_t)vy) * (uint32_t)c.R + (uint32_t)vx;
  return vx >= 0 && vy >= 0 && vz >= 0 && vx < c.R && vy < c.R && vz < c.R && vz >= c.z0 && vz < c.z1;
}

// Shades one fragment and appends it to its voxel's list at arena slot `idx` (reserved by the caller; slots past the capacity are
// dropped and reported by vct_voxelize_stats).  The fragment that finds the voxel empty marks itself in `fresh`: the resolve pass
// and the sparse clear enumerate the occupied voxels through those marks, so no second list (and no second counter) is needed.
__device__ __forceinline__ void push_fragment(const FragCtx& c, const VoxTri& v, uint32_t ti, int i, int j, const float b[3], F3 pos, uint32_t voxel,
                                              uint32_t idx) {
  if (idx >= c.frag_capacity) return;
  FragRec r;
  shade_fragment(v, b, c.mats, c.L, c.cube_size, pos, r.val);
  if (c.accum) {
    // order-independent integer accumulation: the first fragment to arrive claims the voxel (the grid word holds its slot until the resolve
    // pass), every fragment adds its rounded colour to that slot's accumulators.  24-bit sums, 16-bit count: exact up to 65535 fragments.
    const uint32_t prev = atomicCAS(&c.base[(size_t)voxel * c.vstride], 0u, idx + 1u);
    const uint32_t owner = prev ? prev - 1u : idx;
    c.fresh[idx] = prev == 0u ? 1 : 0;
    if (prev == 0u) c.frags[idx].voxel = voxel;
    const unsigned long long q0 = (unsigned long long)(uint32_t)(r.val[0] + 0.5f), q1 = (unsigned long long)(uint32_t)(r.val[1] + 0.5f);
    const unsigned long long q2 = (unsigned long long)(uint32_t)(r.val[2] + 0.5f), q3 = (unsigned long long)(uint32_t)(r.val[3] + 0.5f);
    atomicAdd(c.accum + 2 * (size_t)owner, q0 | (q1 << 24) | (1ull << 48));
    atomicAdd(c.accum + 2 * (size_t)owner + 1, q2 | (q3 << 24));
    return;
  }
  r.next = atomicExch(&c.base[voxel], idx + 1u);
  r.voxel = voxel;
  r.key = ((unsigned long long)ti << 24) | ((unsigned long long)j << 12) | (unsigned long long)i;
  c.frags[idx] = r;
  c.fresh[idx] = r.next == 0u ? 1 : 0;
}

// Fragment candidates of a warp, K per lane (covered = the pixel centre is inside the triangle).  Must be called by all 32 lanes:
// the arena slots of all K*32 candidates are reserved with ONE atomicAdd (a counter per fragment group -- 54 k same-address
// atomics per frame at 256^3 -- was what the raster kernel spent its time on).
template <int K>
__device__ __forceinline__ void emit_fragments(const FragCtx& c, const VoxTri& v, uint32_t ti, const int (&pi)[K], const int (&pj)[K], const float (&pb)[K][3],
                                               bool (&covered)[K], int lane) {
  uint32_t voxel[K], mask[K], total = 0;
  F3 pos[K];
#pragma unroll
  for (int k = 0; k < K; k++) {
    voxel[k] = 0; pos[k] = f3(0.f, 0.f, 0.f);
    if (covered[k]) covered[k] = fragment_voxel(c, v, pb[k], pos[k], voxel[k]);
    mask[k] = __ballot_sync(0xffffffffu, covered[k]);
    total += (uint32_t)__popc(mask[k]);
  }
  if (!total) return;
  uint32_t basei = 0;
  if (lane == 0) basei = atomicAdd(&c.counters[CNT_FRAGS], total);
  basei = __shfl_sync(0xffffffffu, basei, 0);
#pragma unroll
  for (int k = 0; k < K; k++) {
    if (covered[k]) push_fragment(c, v, ti, pi[k], pj[k], pb[k], pos[k], voxel[k], basei + (uint32_t)__popc(mask[k] & ((1u << lane) - 1u)));
    basei += (uint32_t)__popc(mask[k]);
  }
}

// triangles whose bounding box holds at most this many pixel centres are rasterised inside the setup kernel (one lane
// per triangle, the warp walks the boxes in lock-step) instead of becoming 8x8 work items: a scene of millions of
// sub-voxel triangles would otherwise spend a whole warp on one or two fragments
constexpr int kSmallPixels = 100;   // swept on the 1 M / 4 M-triangle scenes (tools/small_limit_sweep.py): 36 -> 100 takes 9 % off the voxelization at 1024^3, flat beyond
// ... and those with up to this many (a few 8x8 blocks) by the whole warp inside the setup kernel, one triangle after the other, 8 x 4
// pixels per step: no VoxTri record, no work items, no prefix searches.  In the 1 M / 4 M-triangle scenes nearly every triangle is of
// this size; as 8x8 work items they were 9.4 of the 16 ms of the voxelization at 1024^3.
constexpr int kMidPixels = 1024;

__global__ void __launch_bounds__(kSetupThreads)
vox_setup_kernel(const vct_vertex_t* __restrict__ verts, const uint32_t* __restrict__ indices, const DrawRec* __restrict__ draws,
                 uint32_t n_draws, uint32_t n_tris, float cube_size, int R, int z0, int z1, VoxTri* __restrict__ out, uint32_t* __restrict__ item_local,
                 uint32_t* __restrict__ item_block, const FragCtx ctx, int small_limit, int mid_limit, uint32_t* scan_ticket, uint32_t* scan_total) {
  __shared__ VoxTri stage[kSetupThreads];   // the mid-sized triangles of the block, one slot per lane: the warp rasterises them together
  uint32_t t = blockIdx.x * kSetupThreads + threadIdx.x;
  uint32_t count = 0;
  VoxTri v;
  v.rt.sign = 0; v.rt.imin = 0; v.rt.imax = -1; v.rt.jmin = 0; v.rt.jmax = -1;
  if (t < n_tris) {
    const DrawRec& d = draws[find_draw(t, draws, n_draws)];
    uint32_t first = d.first_index + 3u * (t - d.tri_base);
    const float* m = d.model;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const vct_vertex_t vx = verts[d.vertex_base + indices[first + k]];
      float px = vx.pos[0], py = vx.pos[1], pz = vx.pos[2];
      float wx = ((m[0] * px + m[4] * py) + m[8] * pz) + m[12];   // model * vec4(position, 1)
      float wy = ((m[1] * px + m[5] * py) + m[9] * pz) + m[13];
      float wz = ((m[2] * px + m[6] * py) + m[10] * pz) + m[14];
      v.wp[k][0] = wx / cube_size; v.wp[k][1] = wy / cube_size; v.wp[k][2] = wz / cube_size;
      const float* nm = d.nmat;
      F3 n = f3((nm[0] * vx.norm[0] + nm[3] * vx.norm[1]) + nm[6] * vx.norm[2],
                (nm[1] * vx.norm[0] + nm[4] * vx.norm[1]) + nm[7] * vx.norm[2],
                (nm[2] * vx.norm[0] + nm[5] * vx.norm[1]) + nm[8] * vx.norm[2]);
      n = normalize(n);
      v.nn[k][0] = n.x; v.nn[k][1] = n.y; v.nn[k][2] = n.z;
    }
    // dominant axis of the face normal; strict comparisons, ties fall through to the (x,z) branch
    F3 w0 = f3(v.wp[0][0], v.wp[0][1], v.wp[0][2]);
    F3 c = cross(sub(f3(v.wp[1][0], v.wp[1][1], v.wp[1][2]), w0), sub(f3(v.wp[2][0], v.wp[2][1], v.wp[2][2]), w0));
    float ax = fabsf(c.x), ay = fabsf(c.y), az = fabsf(c.z);
    uint32_t axis = (az > ax && az > ay) ? 0u : ((ax > ay && ax > az) ? 1u : 2u);
    float xw[3], yw[3];
    const float half = (float)(2 * R) * 0.5f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      float a = axis == 1u ? v.wp[k][1] : v.wp[k][0];
      float b = axis == 0u ? v.wp[k][1] : v.wp[k][2];
      xw[k] = (a + 1.0f) * half;
      yw[k] = (b + 1.0f) * half;
    }
    raster_setup(xw, yw, 2 * R, 2 * R, v.rt);
    if (z0 > 0 || z1 < R) {
      // z-slab voxelization (multi-GPU): a fragment's z is a convex combination of the vertex z's, so a triangle whose
      // voxel-z range (widened by one voxel for interpolation rounding) misses [z0,z1) cannot contribute: no items
      const float zmin = fminf(v.wp[0][2], fminf(v.wp[1][2], v.wp[2][2])), zmax = fmaxf(v.wp[0][2], fmaxf(v.wp[1][2], v.wp[2][2]));
      const float vz_lo = (float)R * (0.5f * zmin + 0.5f) - 1.0f, vz_hi = (float)R * (0.5f * zmax + 0.5f) + 1.0f;
      if (vz_hi < (float)z0 || vz_lo >= (float)z1) { v.rt.sign = 0; v.rt.imin = 0; v.rt.imax = -1; v.rt.jmin = 0; v.rt.jmax = -1; }
    }
    v.material = d.material;
    v.axis = axis;
    count = raster_item_count(v.rt);
  }
  // ---- small triangles: rasterised here ----
  const int bw = v.rt.imax - v.rt.imin + 1, bh = v.rt.jmax - v.rt.jmin + 1;
  const bool small = count > 0 && bw * bh <= small_limit;
  const int npx = small ? bw * bh : 0;
  const bool mid = count > 0 && !small && bw * bh <= mid_limit;
  // The block's item scan comes BEFORE the in-line rasterisation: its barriers are then reached by warps that have all done the same
  // (uniform) set-up work, and the divergent pixel loops below end without anybody waiting for the block's slowest warp.
  {
    uint32_t items = count;
    if (small || mid) items = 0;
    else if (t < n_tris) {
      if (mid_limit > 0 && count) { raster_choose_macro(v.rt, kMaxItemsManyTris); items = raster_item_count(v.rt); }
      out[t] = v;   // only triangles that become work items are read again
    }
    block_scan_items(items, t, n_tris, item_local, item_block, scan_ticket, scan_total);
  }
  // pass 1: every lane counts the fragments of its own triangle; pass 2: it writes them at consecutive slots of the warp's
  // reservation -- ONE arena atomic per warp of 32 triangles (one per pixel step of the lock-stepped walk before: 9 M same-address
  // atomics per frame on the 4 M-triangle scene)
  const int lane = threadIdx.x & 31;
  uint32_t mine = 0;
  EdgeBlock eb0;   // edge functions at the box origin (same integers as raster_sample, two multiply-adds per edge and pixel)
  edge_block_setup(v.rt, v.rt.imin, v.rt.jmin, eb0);
  for (int p = 0, ox = 0, oy = 0; p < npx; p++) {   // one flat loop (lanes differ in box shape, not only in size); the column / row counters replace p % bw, p / bw
    float b[3];
    F3 pos; uint32_t voxel;
    if (edge_block_sample(eb0, ox, oy, b) && fragment_voxel(ctx, v, b, pos, voxel)) mine++;
    if (++ox == bw) { ox = 0; oy++; }
  }
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  const uint32_t warp_total = __shfl_sync(0xffffffffu, incl, 31);
  uint32_t slot = 0;
  if (warp_total) {
    if (lane == 0) slot = atomicAdd(&ctx.counters[CNT_FRAGS], warp_total);
    slot = __shfl_sync(0xffffffffu, slot, 0) + (incl - mine);
  }
  for (int p = 0, ox = 0, oy = 0; p < npx; p++) {
    float b[3];
    F3 pos; uint32_t voxel;
    if (edge_block_sample(eb0, ox, oy, b) && fragment_voxel(ctx, v, b, pos, voxel)) push_fragment(ctx, v, t, v.rt.imin + ox, v.rt.jmin + oy, b, pos, voxel, slot++);
    if (++ox == bw) { ox = 0; oy++; }
  }
  // ---- mid-sized triangles: the whole warp on one triangle at a time, in TWO passes over the warp's mid triangles: count the fragments,
  // reserve their arena slots with ONE atomic, write them.  (One reservation per 8 x 4-pixel step -- four or five per triangle -- made the
  // warp wait for an L2 atomic round trip every few dozen instructions: the set-up kernel of the 4 M-triangle scene spent its time there.)
  const uint32_t mid_mask = __ballot_sync(0xffffffffu, mid);
  if (mid_mask) {
    VoxTri* wst = stage + (threadIdx.x & ~31);   // this warp's 32 slots
    if (mid) wst[lane] = v;
    __syncwarp();
    const int lx = lane & 7, ly = lane >> 3;
    uint32_t total = 0;
    for (uint32_t m = mid_mask; m; m &= m - 1u) {
      const VoxTri& sv = wst[__ffs((int)m) - 1];
      const RasterTri& rt = sv.rt;
      EdgeBlock eb;
      edge_block_setup(rt, rt.imin, rt.jmin, eb);
      for (int j0 = rt.jmin; j0 <= rt.jmax; j0 += 4)
        for (int i0 = rt.imin; i0 <= rt.imax; i0 += 8) {   // uniform trip counts (warp-wide ballots)
          const int pi = i0 + lx, pj = j0 + ly;
          float pb[3];
          F3 pos; uint32_t voxel;
          const bool covered = pi <= rt.imax && pj <= rt.jmax && edge_block_sample(eb, pi - rt.imin, pj - rt.jmin, pb) && fragment_voxel(ctx, sv, pb, pos, voxel);
          total += (uint32_t)__popc(__ballot_sync(0xffffffffu, covered));
        }
    }
    if (total) {
      uint32_t basei = 0;
      if (lane == 0) basei = atomicAdd(&ctx.counters[CNT_FRAGS], total);
      basei = __shfl_sync(0xffffffffu, basei, 0);
      for (uint32_t m = mid_mask; m; m &= m - 1u) {
        const int src = __ffs((int)m) - 1;
        const VoxTri& sv = wst[src];
        const uint32_t ti = __shfl_sync(0xffffffffu, t, src);
        const RasterTri& rt = sv.rt;
        EdgeBlock eb;
        edge_block_setup(rt, rt.imin, rt.jmin, eb);
        for (int j0 = rt.jmin; j0 <= rt.jmax; j0 += 4)
          for (int i0 = rt.imin; i0 <= rt.imax; i0 += 8) {
            const int pi = i0 + lx, pj = j0 + ly;
            float pb[3];
            F3 pos; uint32_t voxel;
            const bool covered = pi <= rt.imax && pj <= rt.jmax && edge_block_sample(eb, pi - rt.imin, pj - rt.jmin, pb) && fragment_voxel(ctx, sv, pb, pos, voxel);
            const uint32_t cm = __ballot_sync(0xffffffffu, covered);
            if (covered) push_fragment(ctx, sv, ti, pi, pj, pb, pos, voxel, basei + (uint32_t)__popc(cm & ((1u << lane) - 1u)));
            basei += (uint32_t)__popc(cm);
          }
      }
    }
  }
}

__global__ void __launch_bounds__(256)
vox_raster_kernel(const VoxTri* __restrict__ tris, uint32_t n_tris, const uint32_t* __restrict__ item_local,
                  const uint32_t* __restrict__ item_block, uint32_t n_blocks, const FragCtx ctx) {
  const uint32_t total = ctx.counters[CNT_ITEMS];
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
  // one warp per macro tile (8x8 .. 64x64 pixels, per triangle) (grid-stride): every lane runs the same two binary searches (broadcast loads), the warp rejects
  // the 8x8 blocks no edge reaches and rasterises the rest, two pixels per lane
  for (uint32_t g = warp; g < total; g += n_warps) {
    uint32_t rank;
    const uint32_t ti = find_item_triangle(g, item_block, n_blocks, item_local, n_tris, rank);
    const VoxTri& v = tris[ti];
    const RasterTri rt = v.rt;
    MacroItem mi;
    macro_item_setup(rt, rank, lane, mi);
    for (unsigned long long live = mi.live; live; live &= live - 1ull) {
      const int b = __ffsll((long long)live) - 1;
      EdgeBlock eb;
      macro_block_edges(mi, b, eb);
      const int bx0 = mi.x0 + 8 * (b & 7), by0 = mi.y0 + 8 * (b >> 3);
      int pi[2], pj[2];
      float pb[2][3];
      bool covered[2];
#pragma unroll
      for (int h = 0; h < 2; h++) {   // the 64 pixels of the block: two per lane
        const int p = lane + 32 * h;
        pi[h] = bx0 + (p & 7); pj[h] = by0 + (p >> 3);
        covered[h] = pi[h] >= rt.imin && pi[h] <= rt.imax && pj[h] >= rt.jmin && pj[h] <= rt.jmax && edge_block_sample(eb, p & 7, p >> 3, pb[h]);
      }
      emit_fragments<2>(ctx, v, ti, pi, pj, pb, covered, lane);
    }
  }
}

// ---- imageAtomicRGBA8Avg (voxelize.frag:66-120) applied sequentially ----
__device__ __forceinline__ uint32_t conv_rgba8(const float v[4]) {
  return (((uint32_t)v[3]) & 0xFFu) << 24 | (((uint32_t)v[2]) & 0xFFu) << 16 | (((uint32_t)v[1]) & 0xFFu) << 8 | (((uint32_t)v[0]) & 0xFFu);
}
__device__ __forceinline__ uint32_t enc_nibble(uint32_t m, uint32_t n) {
  return (m & 0xFEFEFEFEu) | (n & 1u) | (n & 2u) << 7 | (n & 4u) << 14 | (n & 8u) << 21;
}
__device__ __forceinline__ uint32_t dec_nibble(uint32_t m) {
  return (m & 1u) | (m & 0x100u) >> 7 | (m & 0x10000u) >> 14 | (m & 0x1000000u) >> 21;
}
__device__ __forceinline__ uint32_t avg_fold(uint32_t stored, const float val[4]) {
  if (stored == 0u) return enc_nibble(conv_rgba8(val), 1u);  // first CAS (expected 0) succeeds
  const uint32_t c = stored & 0xFEFEFEFEu;
  float r[4] = {(float)(c & 0xFFu), (float)((c >> 8) & 0xFFu), (float)((c >> 16) & 0xFFu), (float)((c >> 24) & 0xFFu)};
  uint32_t n = dec_nibble(stored);
  const float fn = (float)n;
  n = n + 1u;
  const float fn1 = (float)n;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    float t = r[k] * fn + val[k];
    t = t / fn1;
    r[k] = rintf(t / 2.0f) * 2.0f;
  }
  return enc_nibble(conv_rgba8(r), n);
}

constexpr int kSortMax = 24;

// 32x8x8 warp-tile (the unit of the streaming mip kernel, mipmap.cu) of a level-0 voxel index; R = 2^logR >= 32
__device__ __forceinline__ uint32_t tile_of_voxel(uint32_t voxel, int logR) {
  const uint32_t m = (1u << logR) - 1u, x = voxel & m, y = (voxel >> logR) & m, z = voxel >> (2 * logR);
  return ((((z >> 3) << (logR - 3)) + (y >> 3)) << (logR - 5)) + (x >> 5);
}

// vct_grid_clear, sparse form: zero the voxels (and the tile flags) the last voxelization occupied = the voxels of its `fresh` fragments
__global__ void __launch_bounds__(256)
sparse_clear_kernel(uint32_t* __restrict__ base, const FragRec* __restrict__ frags, const uint8_t* __restrict__ fresh, const uint32_t* __restrict__ counters,
                    uint32_t capacity, uint8_t* __restrict__ tile_touched, int logR) {
  const uint32_t n = min(counters[CNT_FRAGS], capacity);
  for (uint32_t f = blockIdx.x * blockDim.x + threadIdx.x; f < n; f += gridDim.x * blockDim.x) {
    if (!fresh[f]) continue;
    const uint32_t voxel = frags[f].voxel;
    base[voxel] = 0u;
    if (tile_touched) tile_touched[tile_of_voxel(voxel, logR)] = 0;
  }
}

// Resolves the voxel whose first fragment sits in arena slot f: sorts the voxel's fragment list by the canonical key and folds it with the
// reference's running average (or reads the fixed-point accumulators), stores the word, marks the mip tile, multi-GPU: stores into the
// peers.  Returns the number of fragments of the voxel.
__device__ __forceinline__ uint32_t resolve_voxel(uint32_t f, uint32_t list_pos, uint32_t* __restrict__ base, const FragRec* __restrict__ frags, const PeerView& pv,
                                                  uint8_t* __restrict__ tile_touched, int logR, unsigned long long* __restrict__ accum, int fmt16) {
  const uint32_t voxel = frags[f].voxel;
  uint32_t stored = 0u, n = 0;
  if (accum) {
    // fixed-point variant: rounded integer mean of the voxel's fragments; the accumulators are left zero for the next frame
    const unsigned long long a0 = accum[2 * (size_t)f], a1 = accum[2 * (size_t)f + 1];
    accum[2 * (size_t)f] = 0ull; accum[2 * (size_t)f + 1] = 0ull;
    n = (uint32_t)(a0 >> 48);
    const uint32_t h = n >> 1, s0 = (uint32_t)(a0 & 0xFFFFFFu), s1 = (uint32_t)((a0 >> 24) & 0xFFFFFFu), s2 = (uint32_t)(a1 & 0xFFFFFFu), s3 = (uint32_t)((a1 >> 24) & 0xFFFFFFu);
    stored = ((s0 + h) / n) | ((s1 + h) / n) << 8 | ((s2 + h) / n) << 16 | ((s3 + h) / n) << 24;
    if (fmt16) {
      // RGBA16F storage variant: the mean colour in [0,1] rounded to half (same expression as the oracle, IEEE division)
      const float dn = (float)n * 255.0f;
      const unsigned long long h0 = __half_as_ushort(__float2half_rn((float)s0 / dn)), h1 = __half_as_ushort(__float2half_rn((float)s1 / dn));
      const unsigned long long h2 = __half_as_ushort(__float2half_rn((float)s2 / dn)), h3 = __half_as_ushort(__float2half_rn((float)s3 / dn));
      reinterpret_cast<unsigned long long*>(base)[voxel] = h0 | (h1 << 16) | (h2 << 32) | (h3 << 48);
      return n;
    }
  } else {
    const uint32_t head = base[voxel];
    unsigned long long keys[kSortMax];
    uint32_t ids[kSortMax];
    for (uint32_t node = head; node != 0u; node = frags[node - 1u].next) {
      if (n < kSortMax) {
        // insertion sort, ascending key
        unsigned long long k = frags[node - 1u].key;
        int pos = (int)n;
        while (pos > 0 && keys[pos - 1] > k) { keys[pos] = keys[pos - 1]; ids[pos] = ids[pos - 1]; pos--; }
        keys[pos] = k; ids[pos] = node - 1u;
      }
      n++;
    }
    if (n <= (uint32_t)kSortMax) {
      for (uint32_t i = 0; i < n; i++) stored = avg_fold(stored, frags[ids[i]].val);
    } else {
      // long list: repeated selection of the next key (O(n^2) list walks, no extra storage)
      unsigned long long last = 0ull;
      bool first = true;
      for (uint32_t i = 0; i < n; i++) {
        unsigned long long best = ~0ull;
        uint32_t best_id = 0u;
        for (uint32_t node = head; node != 0u; node = frags[node - 1u].next) {
          unsigned long long k = frags[node - 1u].key;
          if ((first || k > last) && k <= best) { best = k; best_id = node - 1u; }
        }
        stored = avg_fold(stored, frags[best_id].val);
        last = best;
        first = false;
      }
    }
  }
  base[voxel] = stored;
  const uint32_t tile = tile_of_voxel(voxel, logR);
  if (tile_touched) tile_touched[tile] = 1;   // sparse mip build: this 32x8x8 tile has content
  // multi-GPU: the slab owner writes the resolved voxel straight into every peer's grid over NVLink (sparse
  // exchange: only occupied voxels travel; replaces the dense all-gather of the base level)
  for (int p = 0; p < pv.nranks; p++) {
    if (p != pv.rank) pv.base[p][voxel] = stored;
    if (pv.touched[p]) pv.touched[p][tile] = 1;   // keeps every rank's mip build sparse (own rank included)
  }
  if (pv.pushed && list_pos < pv.pushed_capacity) pv.pushed[list_pos] = voxel;
  return n;
}

// The warp scans arena slots, 32 at a time, and queues the slots of voxels' FIRST fragments (`fresh`); whenever 32 are queued every
// lane resolves one voxel.  (One thread per slot, the fresh ones resolving in place, left a third of the lanes -- one voxel per 3.2
// fragments in the large scenes -- chasing list links while the others idled: the kernel is bound by the latency of those dependent
// loads, and the loads in flight per warp are what it has to offer against it.)
__global__ void __launch_bounds__(128)
vox_resolve_kernel(uint32_t* __restrict__ base, const FragRec* __restrict__ frags, const uint8_t* __restrict__ fresh,
                   uint32_t* __restrict__ counters, uint32_t frag_capacity, const PeerView pv, uint8_t* __restrict__ tile_touched, int logR,
                   uint32_t* __restrict__ status, unsigned long long* __restrict__ accum, int fmt16) {
  __shared__ uint2 queue_s[4][64];   // per warp: (arena slot, position in the pushed list) of the voxels waiting to be resolved
  // arena too small: fragments were dropped.  Tell the host through the mapped status word (the only time this kernel touches host memory)
  if (blockIdx.x == 0 && threadIdx.x == 0 && counters[CNT_FRAGS] > frag_capacity) {
    *reinterpret_cast<volatile uint32_t*>(status + STATUS_OVERFLOW) = counters[CNT_FRAGS];
    __threadfence_system();
  }
  const uint32_t n_frags = min(counters[CNT_FRAGS], frag_capacity);
  uint32_t n_mine = 0, max_list = 0;
  const int lane = threadIdx.x & 31;
  uint2* queue = queue_s[threadIdx.x >> 5];
  int qn = 0;
  for (uint32_t f0 = (blockIdx.x * blockDim.x + threadIdx.x) - lane; f0 < n_frags; f0 += gridDim.x * blockDim.x) {
    const uint32_t f = f0 + lane;
    const bool mine = f < n_frags && fresh[f];
    const uint32_t mm = __ballot_sync(0xffffffffu, mine);
    if (mm) {
      const uint32_t before = (uint32_t)__popc(mm & ((1u << lane) - 1u));
      uint32_t list_pos = 0;
      if (pv.pushed) {   // multi-GPU: the list of pushed voxels is appended with one atomic per warp and round
        if (lane == 0) list_pos = atomicAdd(pv.pushed_n, (uint32_t)__popc(mm));
        list_pos = __shfl_sync(0xffffffffu, list_pos, 0) + before;
      }
      if (mine) queue[qn + (int)before] = make_uint2(f, list_pos);
      qn += __popc(mm);
      __syncwarp();
    }
    if (qn >= 32) {
      qn -= 32;
      const uint2 e = queue[qn + lane];
      __syncwarp();   // the entries are in registers before the next round appends over them
      max_list = max(max_list, resolve_voxel(e.x, e.y, base, frags, pv, tile_touched, logR, accum, fmt16));
      n_mine++;
    }
  }
  if (lane < qn) {
    const uint2 e = queue[lane];
    max_list = max(max_list, resolve_voxel(e.x, e.y, base, frags, pv, tile_touched, logR, accum, fmt16));
    n_mine++;
  }
  // statistics (vct_voxelize_stats): occupied voxels and the longest list, one atomic each per warp
  n_mine = __reduce_add_sync(0xffffffffu, n_mine);
  max_list = __reduce_max_sync(0xffffffffu, max_list);
  if ((threadIdx.x & 31) == 0 && n_mine) { atomicAdd(&counters[CNT_OCCUPIED], n_mine); atomicMax(&counters[CNT_MAXLIST], max_list); }
  if (pv.nranks > 1) peer_signal_last_block(pv, PEER_FLAG_PUSHED, -1);
}

int ensure_tri_scratch(vct_device* dev, int which, size_t n_tris, size_t rec_bytes_total) {
  vct_device::RasterScratch& r = dev->rs[which];
  size_t need = rec_bytes_total;
  if (need > r.tri_recs_bytes) {
    if (r.tri_recs) cudaFree(r.tri_recs);
    r.tri_recs = nullptr; r.tri_recs_bytes = 0;
    size_t want = need + need / 4 + 4096;
    VCT_CUDA(cudaMalloc(&r.tri_recs, want));
    r.tri_recs_bytes = want;
  }
  if (n_tris > r.item_capacity_tris) {
    if (r.item_local) cudaFree(r.item_local);
    if (r.item_block) cudaFree(r.item_block);
    if (r.big_slot) cudaFree(r.big_slot);
    r.item_local = r.item_block = r.big_slot = nullptr; r.item_capacity_tris = 0;
    size_t cap = n_tris + n_tris / 4 + 1024;
    VCT_CUDA(cudaMalloc(&r.item_local, cap * sizeof(uint32_t)));
    if (which == 1) VCT_CUDA(cudaMalloc(&r.big_slot, cap * sizeof(uint32_t)));
    VCT_CUDA(cudaMalloc(&r.item_block, (cap / kSetupThreads + 2) * sizeof(uint32_t)));
    r.item_capacity_tris = cap;
  }
  return VCT_OK;
}

static int log2_int(int v) { int l = 0; while ((1 << l) < v) l++; return l; }

int launch_sparse_clear(vct_device* dev, vct_grid* g) {
  sparse_clear_kernel<<<dev->prop.multiProcessorCount * 4, 256, 0, dev->stream>>>(g->base, dev->frags, dev->fresh, dev->counters, (uint32_t)dev->frag_capacity,
                                                                                   g->tile_touched, log2_int(g->R));
  VCT_CUDA(cudaGetLastError());
  return VCT_OK;
}

int launch_voxelize(vct_device* dev, vct_scene* sc, vct_grid* g, int z0, int z1, const PeerView* push) {
  PeerView pv;
  memset(&pv, 0, sizeof pv);
  if (push) pv = *push;
  if (sc->n_tris == 0 && !push) return VCT_OK;
  int rc = ensure_tri_scratch(dev, 0, sc->n_tris, (size_t)sc->n_tris * sizeof(VoxTri));
  if (rc) return rc;
  if (dev->frag_capacity == 0) {
    rc = vct_voxelize_reserve(dev, 1u << 20);
    if (rc) return rc;
  }
  cudaStream_t s = dev->stream;
  unsigned long long* accum = nullptr;
  if (dev->accum_mode == VCT_ACCUM_FIXED_POINT || g->fmt == VCT_GRID_RGBA16F) {   // (an fp16 grid always accumulates in fixed point)
    if (dev->accum_capacity < dev->frag_capacity) {   // (re)allocated zeroed; the resolve pass zeroes what a frame used
      if (dev->accum) { VCT_CUDA(cudaStreamSynchronize(s)); cudaFree(dev->accum); dev->accum = nullptr; dev->accum_capacity = 0; }
      VCT_CUDA(cudaMalloc(&dev->accum, dev->frag_capacity * 2 * sizeof(unsigned long long)));
      VCT_CUDA(cudaMemsetAsync(dev->accum, 0, dev->frag_capacity * 2 * sizeof(unsigned long long), s));
      dev->accum_capacity = dev->frag_capacity;
    }
    accum = dev->accum;
  }
  // sparse bookkeeping (vct_grid in vct_internal.cuh): the occupied list written below describes level 0 completely only when the
  // level was all zero before; the tile flags stay valid as long as every writer of level 0 marks them
  const bool was_zero = g->base_zero;
  g->base_zero = false;
  if (g->dirty_z0 >= g->dirty_z1) { g->dirty_z0 = z0; g->dirty_z1 = z1; }
  else { g->dirty_z0 = min(g->dirty_z0, z0); g->dirty_z1 = max(g->dirty_z1, z1); }
  g->sparse_clear_ok = was_zero && !push && !g->external;
  dev->vox_owner = g;
  uint8_t* touched = (g->flags_valid && !push && !g->external) ? g->tile_touched : nullptr;
  if (!touched) g->flags_valid = false;
  { int rc2 = launch_fill_u32(s, dev->counters, 8, 0u); if (rc2) return rc2; }   // a kernel, not a memset: no copy-engine work on the frame's critical path
  const uint32_t n_blocks = (sc->n_tris + kSetupThreads - 1) / kSetupThreads;
  VoxTri* tris = (VoxTri*)dev->rs[0].tri_recs;
  const int sms = dev->prop.multiProcessorCount;
  if (sc->n_tris) {   // an empty scene still runs the resolve kernel in multi-GPU mode: the peers wait for its signal
    FragCtx ctx;
    ctx.mats = sc->mats; ctx.L = sc->lights; ctx.cube_size = sc->cube_size; ctx.R = g->R; ctx.z0 = z0; ctx.z1 = z1;
    ctx.base = g->base; ctx.frags = dev->frags; ctx.frag_capacity = (uint32_t)dev->frag_capacity; ctx.fresh = dev->fresh; ctx.counters = dev->counters;
    ctx.accum = accum;
    ctx.vstride = g->fmt == VCT_GRID_RGBA16F ? 2 : 1;
    vox_setup_kernel<<<n_blocks, kSetupThreads, 0, s>>>(sc->verts, sc->indices, sc->draws, sc->n_draws, sc->n_tris, sc->cube_size, g->R, z0, z1, tris,
                                                          dev->rs[0].item_local, dev->rs[0].item_block, ctx, sc->n_tris >= kSmallPathMinTris ? (dev->debug_small_limit >= 0 ? dev->debug_small_limit : kSmallPixels) : 0,
                                                          sc->n_tris >= kSmallPathMinTris ? kMidPixels : 0, dev->counters + CNT_TICKET_VOX, dev->counters + CNT_ITEMS);
    vox_raster_kernel<<<sms * 8, 256, 0, s>>>(tris, sc->n_tris, dev->rs[0].item_local, dev->rs[0].item_block, n_blocks, ctx);
  }
  vox_resolve_kernel<<<sms * 16, 128, 0, s>>>(g->base, dev->frags, dev->fresh, dev->counters, (uint32_t)dev->frag_capacity, pv, touched, log2_int(g->R),
                                             dev->status_dev, accum, g->fmt == VCT_GRID_RGBA16F ? 1 : 0);
  VCT_CUDA(cudaGetLastError());
  return VCT_OK;
}

}  // namespace vct
