// mipmap.cu -- six-direction anisotropic mip chain of the voxel grid.
//
// Replaces Renderer::filter() (src/renderer.cpp:283-314) + shader/mipmap.comp.  The reference dispatches one compute
// pass per level, each re-reading six full textures (and launching 8x more threads than texels).  Here the whole chain
// -- any number of levels -- is built by TWO launches:
//   mip_stream_kernel  persistent warps each take 32x8x8 warp-tiles of level 0 (read ONCE, not 6x, 16-byte loads two slabs ahead of
//                      the arithmetic) -> levels 1, 2, 3 of all six directions and the occupancy bits of levels 0-2.  Pure
//                      streaming: no atomics, no fences, no block-wide barrier, no dependence between warps.  Per round a warp first
//                      compacts the tiles it has to read: with the voxelizer's tile flags, untouched tiles whose outputs are
//                      already zero are neither read nor written.
//   mip_tail_kernel    one wave of independent work on what the fused kernel left in small linear scratch copies:
//                      * one CTA per 64^3 super-block folds its 8^3 level-3 texels into levels 4, 5, 6 (colour + occupancy bits);
//                        the CTA that finishes last builds levels 7.. (a few hundred texels);
//                      * the other CTAs turn the occupancy words of levels 0-2 and the occupancy bytes of level 3 into
//                        the 2x2x2-dilated bit volumes the cone tracer tests, one output word per thread / warp.
// so the dense DRAM traffic is the algorithmic minimum 4 R^3 (read) + 24 R^3 (1/8 + 1/64 + ...) (write) and the traffic of
// a running frame loop is proportional to the occupied tiles.
// Levels >= 1 live ONLY in the stacked mipmapped array the texture units read (surface writes, 16 bytes per store; round 1
// also kept 24-byte records of them: twice the store traffic).
//
// Arithmetic = oracle rules R5/R6, bit-exact: see mip_arith.cuh (integer dot products, ties replayed in fp32; texels with
// many ties fall back to the fp32 recipe).  All-zero child groups are skipped (exact: the filter of zeros is zero);
// mip_generic_kernel / occ_*_kernel cover grids the fused kernel does not (R < 32 or fewer than 6 levels).
#include <cuda_fp16.h>

#include "mip_arith.cuh"
#include "vct_internal.cuh"

namespace vct {

__device__ __forceinline__ void unpack4(uint32_t w, float c[4]) {
  c[0] = unorm8(w, 0); c[1] = unorm8(w, 1); c[2] = unorm8(w, 2); c[3] = unorm8(w, 3);
}

// ---- the fp32 recipe for a whole texel (mipmap.comp:45-100 as the oracle evaluates it) ----
template <int D>
__device__ __forceinline__ uint32_t filter_dir(const float (&c)[8][4]) {
  constexpr uint32_t word = D < 2 ? 0x76543210u : (D < 4 ? 0x67324510u : 0x75316420u);
  constexpr uint32_t fronts = (D & 1) ? word >> 16 : word & 0xFFFFu, backs = (D & 1) ? word & 0xFFFFu : word >> 16;
  float r[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < 4; p++) {
      const int f = (fronts >> (4 * p)) & 7, b = (backs >> (4 * p)) & 7;
      const float v = __fadd_rn(c[f][k], __fmul_rn(__fadd_rn(1.0f, -c[f][3]), c[b][k]));
      s = p == 0 ? v : __fadd_rn(s, v);
    }
    r[k] = __fadd_rn(fminf(__fmul_rn(s, 63.75f), 255.0f), 12582912.0f);   // (s / 4) * 255 rounds once: s / 4 is exact
  }
  const uint32_t lo = __byte_perm(__float_as_uint(r[0]), __float_as_uint(r[1]), 0x0040u);
  const uint32_t hi = __byte_perm(__float_as_uint(r[2]), __float_as_uint(r[3]), 0x0040u);
  return __byte_perm(lo, hi, 0x5410u);
}
__device__ __forceinline__ void filter6_fp32(const uint32_t (&w)[8], uint32_t (&o)[6]) {
  float c[8][4];
#pragma unroll
  for (int i = 0; i < 8; i++) unpack4(w[i], c[i]);
  o[0] = filter_dir<0>(c); o[1] = filter_dir<1>(c); o[2] = filter_dir<2>(c);
  o[3] = filter_dir<3>(c); o[4] = filter_dir<4>(c); o[5] = filter_dir<5>(c);
}

// child i of a 2x2x2 group whose low corner is `p`, rows `sy` and slices `sz` words apart (the replay's run-time index: an
// address, not a chain of selects over the register copy)
struct ChildLoader {
  const uint32_t* p; int sy, sz;
  __device__ __forceinline__ uint32_t operator()(int i) const { return p[(((i & 1) ^ 1) * sz + (((i >> 1) & 1) ^ 1) * sy) + (((i >> 2) & 1) ^ 1)]; }
};

// six directions of one texel whose children are shared by all directions (level 0 -> 1); all 32 lanes call it
__device__ __forceinline__ void filter6_shared(const uint32_t (&w)[8], bool any, const ChildLoader& ld, uint32_t (&o)[6]) {
#pragma unroll
  for (int d = 0; d < 6; d++) o[d] = 0u;
  uint32_t ties = 0u;
  if (any) mip_filter6_shared(w, o, ties);
  // many ties (binary alpha: a quarter of the channels) -> the fp32 recipe for the whole texel is cheaper than replaying them
  if (__any_sync(0xffffffffu, __popc(ties) > 2)) {
    if (ties) filter6_fp32(w, o);
    return;
  }
  while (__any_sync(0xffffffffu, ties != 0u)) {
    if (ties) {
      const int bit = __ffs((int)ties) - 1, d = bit >> 2, k = bit & 3;
      const uint32_t b = mip_replay_channel(ld, d, k);
      const uint32_t m = 0xFFu << (8 * k);
#pragma unroll
      for (int dd = 0; dd < 6; dd++)
        if (dd == d) o[dd] = (o[dd] & ~m) | (b << (8 * k));
      ties &= ties - 1u;
    }
  }
}

// one direction of one texel from its own eight children (level >= 1 -> next); any thread may call it
template <class Load>
__device__ __forceinline__ uint32_t filter1(const uint32_t (&w)[8], int d, const Load& ld) {
  uint32_t any = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) any |= w[i];
  if (!any) return 0u;
  uint32_t ties;
  uint32_t out = mip_filter1(w, d, ties);
  for (; ties; ties &= ties - 1u) {
    const int k = __ffs((int)ties) - 1;
    const uint32_t b = mip_replay_channel(ld, d, k);
    out = (out & ~(0xFFu << (8 * k))) | (b << (8 * k));
  }
  return out;
}
// the same when the children only exist in registers (cold paths: generic kernel, block fold, top of the chain)
__device__ __forceinline__ uint32_t filter1(const uint32_t (&w)[8], int d) {
  return filter1(w, d, [&](int i) {
    uint32_t v = w[0];
#pragma unroll
    for (int j = 1; j < 8; j++) v = i == j ? w[j] : v;
    return v;
  });
}

// One direction of one texel with the fp32 recipe itself (no tie logic, no divergent replay): what the tail kernel uses.  Its
// super-block folds are a few hundred texels on the critical path of the launch, and the faint values of coarse levels of a sparse
// scene sit on rounding ties all the time (N = 510: two children of value 1) -- with the integer path + per-tie replay a fold CTA
// ran 7 k instructions per thread, 13 us of pure latency at 256^3.
__device__ __noinline__ uint32_t filter1_fp32(const uint32_t (&w)[8], int d) {   // one copy: the tail kernel calls it from four places, and its code size is start-up latency (instruction fetch)
  uint32_t any = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) any |= w[i];
  if (!any) return 0u;
  float c[8][4];
#pragma unroll
  for (int i = 0; i < 8; i++) unpack4(w[i], c[i]);
  switch (d) {
    case 0: return filter_dir<0>(c);
    case 1: return filter_dir<1>(c);
    case 2: return filter_dir<2>(c);
    case 3: return filter_dir<3>(c);
    case 4: return filter_dir<4>(c);
    default: return filter_dir<5>(c);
  }
}

// OR of adjacent bit pairs: bit k of the result = bit 2k | bit 2k+1 of the 32-bit input (16 result bits)
__host__ __device__ __forceinline__ uint32_t occ_pair_or(uint32_t v) {
  v = (v | (v >> 1)) & 0x55555555u;
  v = (v | (v >> 1)) & 0x33333333u;
  v = (v | (v >> 2)) & 0x0F0F0F0Fu;
  v = (v | (v >> 4)) & 0x00FF00FFu;
  v = (v | (v >> 8)) & 0x0000FFFFu;
  return v;
}

// ---------------------------------------------------------------------------------------------
// generic fallback, one level per launch: one thread per (destination texel, direction); source = level 0 words or the
// array level below (surface reads: the previous launch wrote it)
__global__ void mip_generic_kernel(const uint32_t* __restrict__ base, cudaSurfaceObject_t src, int src_pitch, cudaSurfaceObject_t dst, int dst_pitch, int Ns, int Nd) {
  const size_t n = (size_t)Nd * Nd * Nd * 6;
  for (size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x; u < n; u += (size_t)gridDim.x * blockDim.x) {
    const int d = (int)(u % 6);
    size_t tex = u / 6;
    const int x = (int)(tex % Nd), y = (int)((tex / Nd) % Nd), z = (int)(tex / ((size_t)Nd * Nd));
    uint32_t w[8];
#pragma unroll
    for (int dz = 0; dz < 2; dz++)
#pragma unroll
      for (int dy = 0; dy < 2; dy++)
#pragma unroll
        for (int dx = 0; dx < 2; dx++)
          w[child_id(dx, dy, dz)] = base ? base[((size_t)(2 * z + dz) * Ns + (2 * y + dy)) * Ns + (2 * x + dx)]
                                         : surf3Dread<uint32_t>(src, (2 * x + dx) * 4, 2 * y + dy, 2 * z + dz + d * src_pitch);
    surf3Dwrite(filter1(w, d), dst, x * 4, y, z + d * dst_pitch);
  }
}

// ---------------------------------------------------------------------------------------------
// The streaming kernel.  Work unit = one WARP-TILE of 32 x 8 x 8 level-0 voxels (8 KB, 64 rows of 128 bytes) that ONE warp
// takes through levels 1, 2 and 3 on its own: no block-wide barrier, no shared state between warps, so a stalled warp (waiting
// for DRAM, replaying a rounding tie) never holds seven others at a __syncthreads the way the CTA-wide 16 KB tiles of the
// TMA-ring version did (1.8 barrier-stalled warps per issue slot, 3500 cycles per tile per CTA even for empty tiles).
//  * Level 0 is read ONCE, with 16-byte asynchronous copies (cp.async, L1 bypassed) into a per-warp ring of three 2 KB slabs in
//    shared memory, two z-slabs ahead of the arithmetic: 12 warps x 4 KB = 48 KB in flight per SM.  (Loads into registers cannot
//    run two slabs ahead in a rolled loop -- rotating the buffers with MOVs waits for the newest load -- and unrolling the slab loop
//    would copy the arithmetic four times.)  Every lane reads back exactly the 64 bytes it copied, so no barrier is involved.
//    Lane mapping of a slab (z1 = 0..3): q = lane & 7 -> voxels x = 4q .. 4q+3 (one 16-byte copy per row), r = lane >> 3 ->
//    level-1 row y1 = r; the lane owns level-1 texels x1 = 2q, 2q+1 of (y1, z1).  A warp-wide copy = four full 128-byte lines.
//  * The NON-ZERO level-1 texels of a slab are appended to a per-warp queue in shared memory (eight children + the texel's
//    slot) and the arithmetic runs on full batches of 32 queue entries: in a real scene a non-empty tile holds a few dozen
//    non-zero texels out of 256 (a surface crossing it), and running the 300-instruction filter with 4 of 32 lanes live
//    was what made one such tile cost 15 us of one warp's time.  Level 2 compacts its (texel, direction) items the same way.
//  * Tiles are handed out dynamically (one device counter): a warp that meets a heavy tile simply takes fewer tiles.
constexpr int WX = 32, WY = 8, WZ = 8;
constexpr int kStreamThreads = 192, kStreamWarps = kStreamThreads / 32;
constexpr int kRingSlabs = 3;
constexpr int kQueueRows = 96;   // at most 31 entries wait when a slab appends up to 64

struct StreamArgs {
  const uint32_t* base;
  int R, levels;
  uint32_t n_tiles;
  int log_tx, log_ty;        // tiles per axis as shifts: R/32 in x, R/8 in y (and z)
  uint32_t* occ0; uint16_t* occ1; uint8_t* occ2;
  uint8_t* tile_zero;        // per tile: 1 = every output of this tile is known to be zero
  const uint8_t* touched;    // per tile: the voxelizer wrote into it since the last clear (nullptr: unknown, every tile is read)
  int dense;                 // measurement switch: every tile is read AND written (nothing is taken as known)
  uint32_t *zero0, *zero1;   // plain / dilated occupancy words of levels 4..: zeroed here, the tail kernel ORs the occupied texels in
  uint32_t zero0_n, zero1_n;
  uint32_t* rec3;            // [N3^3][6]: level 3 once more, linear, for the tail kernel
  uint8_t* occb3;            // [N3^3]: level-3 occupancy bytes
  uint32_t* counter;         // work counter (zero at launch; the tail kernel resets it)
  uint32_t* sb_epoch;        // per 64^3 super-block: number of the last build that processed one of its tiles (the tail kernel skips the others)
  uint32_t build;
  int log_nsb;               // super-blocks per axis as a shift (0 when R <= 64)
  SurfSet surf;
};

struct WarpSmem {
  uint4 ring[kRingSlabs][4][32];            // 6 KB: level-0 slabs in flight / being reduced: [slab][dz * 2 + dy][lane]
  uint32_t s1[6][WZ / 2][WY / 2][WX / 2];   // 6 KB: level 1 of the warp-tile, one block per direction
  uint32_t s2[6][WZ / 4][WY / 4][WX / 4];   // 768 B
  uint32_t s3[6][WX / 8];                   // 96 B
  uint32_t rows[WZ][WY];                    // level-0 occupancy word of every row of the tile
  uint32_t queue[kQueueRows][9];            // non-zero level-1 texels waiting for the arithmetic: children [dz][dy][dx], slot in s1
  uint32_t list[32];                        // ring of tiles to process (bit 31: tile_zero[] of the tile)
  uint32_t tex2[32];                        // the non-zero level-2 texels of the tile
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {   // 16 bytes global -> shared, asynchronous, not kept in L1
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct TileCoord { int x0, y0, z0, tx; };
__device__ __forceinline__ TileCoord tile_coord(uint32_t tile, const StreamArgs& a) {
  TileCoord c;
  c.tx = (int)(tile & ((1u << a.log_tx) - 1u));
  c.x0 = c.tx * WX;
  c.y0 = (int)((tile >> a.log_tx) & ((1u << a.log_ty) - 1u)) * WY;
  c.z0 = (int)(tile >> (a.log_tx + a.log_ty)) * WZ;
  return c;
}
// the four 16-byte copies of one lane for slab z1 of a tile: index dz * 2 + dy
__device__ __forceinline__ void issue_slab(uint4 (*dst)[32], const StreamArgs& a, const TileCoord& c, int z1, int lane, int q, int r) {
  const uint32_t* p = a.base + ((size_t)(c.z0 + 2 * z1) * a.R + (c.y0 + 2 * r)) * a.R + c.x0 + 4 * q;
  const size_t sy = (size_t)a.R, sz = (size_t)a.R * a.R;
  cp_async16(&dst[0][lane], p); cp_async16(&dst[1][lane], p + sy); cp_async16(&dst[2][lane], p + sz); cp_async16(&dst[3][lane], p + sz + sy);
}

// six directions of one level-1 texel from its eight level-0 children (shared by all directions); all 32 lanes call it.
// `row` = the lane's queue entry (children in [dz][dy][dx] order = what ChildLoader{row, 2, 4} reads): ties are replayed from it.
__device__ __forceinline__ void filter6_row(const uint32_t (&w)[8], bool any, const uint32_t* row, uint32_t (&o)[6]) {
#pragma unroll
  for (int d = 0; d < 6; d++) o[d] = 0u;
  uint32_t ties = 0u;
  if (any) mip_filter6_shared(w, o, ties);
  if (!__any_sync(0xffffffffu, ties != 0u)) return;
  // many ties (binary alpha: a quarter of the channels) -> the fp32 recipe for the whole texel is cheaper than replaying them
  if (__any_sync(0xffffffffu, __popc(ties) > 2)) {
    if (ties) filter6_fp32(w, o);
    return;
  }
  const ChildLoader ld{row, 2, 4};
  for (; ties; ties &= ties - 1u) {
    const int bit = __ffs((int)ties) - 1, d = bit >> 2, k = bit & 3;
    const uint32_t b = mip_replay_channel(ld, d, k);
    const uint32_t m = 0xFFu << (8 * k);
#pragma unroll
    for (int dd = 0; dd < 6; dd++)
      if (dd == d) o[dd] = (o[dd] & ~m) | (b << (8 * k));
  }
}

// One slab (two z-slices, all 8 rows) of a tile: occupancy rows of level 0; the slab's non-zero level-1 texels (two candidates per
// lane) go to the queue; the arithmetic runs on full batches of the queue (and on what is left when `flush`).  qn = queue length,
// started = the tile has had a non-zero slab (sm.s1 is zeroed when the first one turns up).  Returns "the slab is non-zero".
__device__ __forceinline__ bool process_slab(WarpSmem& sm, const uint4 (*slab)[32], int z1, int lane, int q, int r, bool flush, int& qn, bool& started) {
  const uint4 v[4] = {slab[0][lane], slab[1][lane], slab[2][lane], slab[3][lane]};
  const uint32_t any_a = ((v[0].x | v[0].y) | (v[1].x | v[1].y)) | ((v[2].x | v[2].y) | (v[3].x | v[3].y));
  const uint32_t any_b = ((v[0].z | v[0].w) | (v[1].z | v[1].w)) | ((v[2].z | v[2].w) | (v[3].z | v[3].w));
  const uint32_t mask_a = __ballot_sync(0xffffffffu, any_a != 0u), mask_b = __ballot_sync(0xffffffffu, any_b != 0u);
  const bool nonzero = (mask_a | mask_b) != 0u;
  if (!nonzero) {
    if (q == 0) {
#pragma unroll
      for (int i = 0; i < 4; i++) sm.rows[2 * z1 + (i >> 1)][2 * r + (i & 1)] = 0u;
    }
  } else {
    if (!started) {   // first non-zero slab of the tile: the texels that never reach the queue must read as zero
      const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
      uint4* p = reinterpret_cast<uint4*>(&sm.s1[0][0][0][0]);
#pragma unroll
      for (int i = 0; i < 12; i++) p[lane + 32 * i] = z4;
      started = true;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      uint32_t bits = (min(v[i].x, 1u) | min(v[i].y, 1u) << 1 | min(v[i].z, 1u) << 2 | min(v[i].w, 1u) << 3) << (4 * q);
      bits |= __shfl_xor_sync(0xffffffffu, bits, 1);
      bits |= __shfl_xor_sync(0xffffffffu, bits, 2);
      bits |= __shfl_xor_sync(0xffffffffu, bits, 4);
      if (q == 0) sm.rows[2 * z1 + (i >> 1)][2 * r + (i & 1)] = bits;
    }
    const uint32_t lt = (1u << lane) - 1u;
    const int na = __popc(mask_a);
    const uint32_t slot = (uint32_t)((z1 * 4 + r) * 16 + 2 * q);
    if (any_a) {
      uint32_t* row = sm.queue[qn + __popc(mask_a & lt)];
#pragma unroll
      for (int i = 0; i < 4; i++) { row[2 * i] = v[i].x; row[2 * i + 1] = v[i].y; }
      row[8] = slot;
    }
    if (any_b) {
      uint32_t* row = sm.queue[qn + na + __popc(mask_b & lt)];
#pragma unroll
      for (int i = 0; i < 4; i++) { row[2 * i] = v[i].z; row[2 * i + 1] = v[i].w; }
      row[8] = slot + 1u;
    }
    qn += na + __popc(mask_b);
    __syncwarp();
  }
  // ---- the arithmetic: batches of 32 entries from the end of the queue ----
  while (qn >= 32 || (flush && qn > 0)) {
    const int nb = min(qn, 32), e0 = qn - nb;
    const bool live = lane < nb;
    const uint32_t* row = sm.queue[e0 + (live ? lane : 0)];
    uint32_t w[8];
#pragma unroll
    for (int dz = 0; dz < 2; dz++)
#pragma unroll
      for (int dy = 0; dy < 2; dy++)
#pragma unroll
        for (int dx = 0; dx < 2; dx++) w[child_id(dx, dy, dz)] = live ? row[dz * 4 + dy * 2 + dx] : 0u;
    const uint32_t slot = row[8];
    uint32_t o[6];
    filter6_row(w, live, row, o);
    if (live) {
      uint32_t* dst = &sm.s1[0][0][0][0] + slot;
#pragma unroll
      for (int d = 0; d < 6; d++) dst[d * (WZ / 2) * (WY / 2) * (WX / 2)] = o[d];
    }
    qn = e0;
    __syncwarp();   // the rows of this batch may be overwritten by the next append
  }
  return nonzero;
}

// levels 1-3 and the occupancy words of an all-zero tile: zero stores, in four parts.  The parts of tile i are issued between the
// slab steps of tile i + 1 (or all at once when there is none): a surface store holds the issuing warp until the texture unit has
// taken its operands, and twelve of them back to back at the end of a tile left the tile's two prefetched slabs waiting.
__device__ __forceinline__ void store_zero_part(const StreamArgs& a, const TileCoord& c, int lane, int part) {
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  const int R = a.R, N1 = R >> 1, N2 = R >> 2, N3 = R >> 3;
  {
    // level 1: the twelve stores of a lane differ only in z (slice within the tile, direction); three per part
    const int xb = (c.x0 / 2 + 4 * (lane & 3)) * 4, y = c.y0 / 2 + ((lane >> 2) & 3), zb = c.z0 / 2 + (lane >> 4);
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const int i = 3 * part + j;
      surf3Dwrite(zero, a.surf.s[1], xb, y, zb + 2 * (i & 1) + (i >> 1) * a.surf.pitch[1]);
    }
  }
  if (part == 0) {
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const int u = lane + 32 * i;
      if (u < 48) surf_write(a.surf, u >> 3, 2, zero, (c.x0 / 4 + 4 * (u & 1)) * 4, c.y0 / 4 + ((u >> 1) & 1), c.z0 / 4 + ((u >> 2) & 1));
    }
  } else if (part == 1) {
    const size_t t3 = ((size_t)(c.z0 / 8) * N3 + c.y0 / 8) * N3 + c.x0 / 8;
    if (lane < 6) surf_write(a.surf, lane, 3, zero, (c.x0 / 8) * 4, c.y0 / 8, c.z0 / 8);
    if (lane < 24) __stcg(a.rec3 + t3 * 6 + lane, 0u);
    if (lane >= 24 && lane < 28) __stcg(a.occb3 + t3 + (lane - 24), (uint8_t)0);
  } else if (part == 2) {
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const int z = (lane >> 3) + 4 * i, y = lane & 7;
      a.occ0[((size_t)(c.z0 + z) * R + (c.y0 + y)) * (R / 32) + c.tx] = 0u;
    }
  } else {
    if (lane < 16) a.occ1[((size_t)(c.z0 / 2 + (lane >> 2)) * N1 + (c.y0 / 2 + (lane & 3))) * (N1 / 16) + c.tx] = (uint16_t)0;
    if (lane >= 16 && lane < 20) a.occ2[((size_t)(c.z0 / 4 + ((lane - 16) >> 1)) * N2 + (c.y0 / 4 + (lane & 1))) * (N2 / 8) + c.tx] = (uint8_t)0;
  }
}

// the rest of a non-zero tile once its four slabs are in sm.s1 / sm.rows: level-1 stores, levels 2 and 3, occupancy words
__device__ __forceinline__ void finish_tile(WarpSmem& sm, const StreamArgs& a, const TileCoord& c, int lane) {
  const int R = a.R, N1 = R >> 1, N2 = R >> 2, N3 = R >> 3;
  __syncwarp();
  // ---- level 1 -> the array: 6 x 4 x 4 rows of 64 bytes, 16 bytes per lane and store ----
  {
    const int xq = lane & 3, yl = (lane >> 2) & 3, zl = lane >> 4;
    const int xb = (c.x0 / 2 + 4 * xq) * 4, y = c.y0 / 2 + yl, zb = c.z0 / 2 + zl;
    uint4 v[12];
#pragma unroll
    for (int i = 0; i < 12; i++) v[i] = *reinterpret_cast<const uint4*>(&sm.s1[i >> 1][zl + 2 * (i & 1)][yl][4 * xq]);
#pragma unroll
    for (int i = 0; i < 12; i++) surf3Dwrite(v[i], a.surf.s[1], xb, y, zb + 2 * (i & 1) + (i >> 1) * a.surf.pitch[1]);
  }
  // ---- level 2: the texels with a non-zero level-0 support (one candidate per lane) x 6 directions, compacted ----
  {
    const int x = lane & 7, y = (lane >> 3) & 1, z = lane >> 4;
    uint32_t sup = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) sup |= sm.rows[4 * z + (k >> 2)][4 * y + (k & 3)];
    const bool occupied = ((sup >> (4 * x)) & 0xFu) != 0u;
    const uint32_t m2 = __ballot_sync(0xffffffffu, occupied);
    if (occupied) sm.tex2[__popc(m2 & ((1u << lane) - 1u))] = (uint32_t)lane;
#pragma unroll
    for (int d = 0; d < 6; d++) sm.s2[d][z][y][x] = 0u;
    __syncwarp();
    const int n2 = __popc(m2), n_items = 6 * n2;
#pragma unroll 1
    for (int i0 = 0; i0 < n_items; i0 += 32) {
      const int i = i0 + lane;
      if (i < n_items) {
        const int d = i / n2, t = (int)sm.tex2[i - d * n2], tx = t & 7, ty = (t >> 3) & 1, tz = t >> 4;
        uint32_t w[8];
#pragma unroll
        for (int dz = 0; dz < 2; dz++)
#pragma unroll
          for (int dy = 0; dy < 2; dy++) {
            const uint2 p = *reinterpret_cast<const uint2*>(&sm.s1[d][2 * tz + dz][2 * ty + dy][2 * tx]);
            w[child_id(0, dy, dz)] = p.x;
            w[child_id(1, dy, dz)] = p.y;
          }
        const ChildLoader ld{&sm.s1[d][2 * tz][2 * ty][2 * tx], WX / 2, (WX / 2) * (WY / 2)};
        sm.s2[d][tz][ty][tx] = filter1(w, d, ld);
      }
    }
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const int u = lane + 32 * i;
    if (u < 48) {
      const int xq = u & 1, y = (u >> 1) & 1, z = (u >> 2) & 1, d = u >> 3;
      const uint4 v = *reinterpret_cast<const uint4*>(&sm.s2[d][z][y][4 * xq]);
      surf_write(a.surf, d, 2, v, (c.x0 / 4 + 4 * xq) * 4, c.y0 / 4 + y, c.z0 / 4 + z);
    }
  }
  // ---- level 3: 4 texels x 6 directions ----
  const size_t t3 = ((size_t)(c.z0 / 8) * N3 + c.y0 / 8) * N3 + c.x0 / 8;
  if (lane < 24) {
    const int d = lane >> 2, x = lane & 3;
    uint32_t w[8];
#pragma unroll
    for (int dz = 0; dz < 2; dz++)
#pragma unroll
      for (int dy = 0; dy < 2; dy++)
#pragma unroll
        for (int dx = 0; dx < 2; dx++) w[child_id(dx, dy, dz)] = sm.s2[d][dz][dy][2 * x + dx];
    const ChildLoader ld{&sm.s2[d][0][0][2 * x], WX / 4, (WX / 4) * (WY / 4)};
    const uint32_t o = filter1(w, d, ld);
    sm.s3[d][x] = o;
    __stcg(a.rec3 + (t3 + x) * 6 + d, o);   // the linear copy the tail kernel folds into levels 4..
  }
  __syncwarp();
  if (lane < 6) surf_write(a.surf, lane, 3, *reinterpret_cast<const uint4*>(&sm.s3[lane][0]), (c.x0 / 8) * 4, c.y0 / 8, c.z0 / 8);
  // ---- occupancy words.  A texel of level >= 1 is occupied when a voxel of its level-0 support is non-zero (a superset of "the
  // texel is non-zero": a filtered value can round to zero), so the bits of a level are ORs of the level-0 row words ----
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const int z = (lane >> 3) + 4 * i, y = lane & 7;
    a.occ0[((size_t)(c.z0 + z) * R + (c.y0 + y)) * (R / 32) + c.tx] = sm.rows[z][y];
  }
  if (lane < 16) {   // level 1: 16 rows of 16 bits
    const int y = lane & 3, z = lane >> 2;
    const uint32_t w = (sm.rows[2 * z][2 * y] | sm.rows[2 * z][2 * y + 1]) | (sm.rows[2 * z + 1][2 * y] | sm.rows[2 * z + 1][2 * y + 1]);
    a.occ1[((size_t)(c.z0 / 2 + z) * N1 + (c.y0 / 2 + y)) * (N1 / 16) + c.tx] = (uint16_t)occ_pair_or(w);
  } else if (lane < 20) {   // level 2: 4 rows of 8 bits
    const int y = lane & 1, z = (lane - 16) >> 1;
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) w |= sm.rows[4 * z + (k >> 2)][4 * y + (k & 3)];
    a.occ2[((size_t)(c.z0 / 4 + z) * N2 + (c.y0 / 4 + y)) * (N2 / 8) + c.tx] = (uint8_t)occ_pair_or(occ_pair_or(w));
  } else if (lane == 24) {   // level 3: 4 texels, one occupancy BYTE each (sub-byte bit rows would be shared between tiles)
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < 64; k++) w |= sm.rows[k >> 3][k & 7];
#pragma unroll
    for (int x = 0; x < 4; x++) __stcg(a.occb3 + t3 + x, (uint8_t)(((w >> (8 * x)) & 0xFFu) != 0u));
  }
  __syncwarp();   // sm.s1 / sm.rows are rewritten by the next tile
}

// Work distribution.  The tiles a warp has to process sit in a small ring in shared memory (sm.list); loads run two slabs ahead
// of the arithmetic and cross from one tile of the ring into the next.
//   dense (nothing known about level 0: every tile is read): the first two tiles of a warp are static, every further one comes
//     from the device counter; the request goes out at the end of a tile and is collected at the end of the next one, so its
//     latency is covered and the ring always holds the tile being processed plus the next;
//   sparse (voxelizer tile flags): the counter hands out CHUNKS of 8 candidate tiles scattered over the grid (chunk_tile(): the occupied
//     tiles of a scene are neighbours -- a wall is a plane of them -- and would otherwise land on one warp); the warp keeps the
//     touched tiles and those whose outputs of the previous build are not zero yet (< 5 % of the tiles of the Cornell scene).
constexpr int kChunkTiles = 8;
// candidate j of chunk c: the j-th eighth of the tile range, position scrambled per eighth (a bijection: n_chunks is a power of two)
__device__ __forceinline__ uint32_t chunk_tile(uint32_t c, int j, uint32_t n_chunks) {
  return (uint32_t)j * n_chunks + ((c ^ ((uint32_t)j * 0x9E5u)) & (n_chunks - 1u));
}
__global__ void __launch_bounds__(kStreamThreads, 2)
mip_stream_kernel(const StreamArgs a) {
  extern __shared__ __align__(16) unsigned char stream_smem_raw[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, q = lane & 7, r = lane >> 3;
  WarpSmem& sm = reinterpret_cast<WarpSmem*>(stream_smem_raw)[wid];
  for (uint32_t i = blockIdx.x * kStreamThreads + threadIdx.x; i < a.zero0_n; i += gridDim.x * kStreamThreads) a.zero0[i] = 0u;
  for (uint32_t i = blockIdx.x * kStreamThreads + threadIdx.x; i < a.zero1_n; i += gridDim.x * kStreamThreads) a.zero1[i] = 0u;
  const uint32_t gwarp = blockIdx.x * kStreamWarps + wid, n_warps = gridDim.x * kStreamWarps;
  const bool sparse = a.touched != nullptr;
  const uint32_t n_chunks = a.n_tiles / kChunkTiles;   // n_tiles = R^3 / 2048 is a power of two >= 16
  uint32_t head = 0, tail = 0;        // ring entries [head, tail), index & 31
  uint32_t pend = 0;                  // dense: the tile requested from the counter (lane 0)
  bool pending = false, exhausted = false;
  uint32_t next_chunk = gwarp;        // sparse: the chunk to examine next
  if (!sparse) {
    if (gwarp < a.n_tiles) { if (lane == 0) sm.list[0] = gwarp; tail = 1; }
    if (gwarp + n_warps < a.n_tiles) { if (lane == 0) sm.list[1] = gwarp + n_warps; tail = 2; }
    else exhausted = true;
    if (!exhausted) {
      if (lane == 0) pend = 2u * n_warps + atomicAdd(a.counter, 1u);
      pending = true;
    }
    __syncwarp();
  }
  uint32_t sc = 0;                    // slabs consumed so far: slab s of the warp's stream lives in ring slot s % 3
  int ahead = 0;                      // slabs of the tile at `head` already requested (0 or 2)
  int qn = 0;
  bool zero_pending = false;          // the zero stores of the previous (empty) tile are still to be issued
  TileCoord cz{};
  while (true) {
    if (head == tail) {
      if (zero_pending) {
#pragma unroll 1
        for (int part = 0; part < 4; part++) store_zero_part(a, cz, lane, part);
        zero_pending = false;
      }
      if (!sparse) break;             // dense: the ring only runs dry when the counter is exhausted
      // ---- sparse: examine chunks until one has work ----
      bool found = false;
      while (!found && next_chunk < n_chunks) {
        uint32_t nc = 0;
        if (lane == 0) nc = n_warps + atomicAdd(a.counter, 1u);   // the chunk after this one: in flight beside the flag loads
        const uint32_t cand = chunk_tile(next_chunk, lane, n_chunks);
        bool active = false;
        uint32_t zero = 0u;
        if (lane < kChunkTiles && cand < a.n_tiles) {
          zero = a.dense ? 0u : a.tile_zero[cand];
          active = a.touched[cand] != 0 || zero == 0u;
        }
        const uint32_t amask = __ballot_sync(0xffffffffu, active);
        if (active) sm.list[(tail + __popc(amask & ((1u << lane) - 1u))) & 31u] = cand | (zero ? 0x80000000u : 0u);
        tail += (uint32_t)__popc(amask);
        found = amask != 0u;
        next_chunk = __shfl_sync(0xffffffffu, nc, 0);
        __syncwarp();
      }
      if (!found) break;
      ahead = 0;
    }
    // ---- one tile ----
    const uint32_t entry = sm.list[head & 31u];
    const uint32_t tile = entry & 0x7FFFFFFFu;
    const TileCoord c = tile_coord(tile, a);
    if (lane == 0) a.sb_epoch[((((uint32_t)c.z0 >> 6) << a.log_nsb) + ((uint32_t)c.y0 >> 6) << a.log_nsb) + ((uint32_t)c.x0 >> 6)] = a.build;
    if (ahead == 0) {
      issue_slab(sm.ring[sc % kRingSlabs], a, c, 0, lane, q, r); cp_async_commit();
      issue_slab(sm.ring[(sc + 1) % kRingSlabs], a, c, 1, lane, q, r); cp_async_commit();
    }
    ahead = 0;
    const bool have_next = head + 1u != tail;
    TileCoord cn = c;
    if (have_next) cn = tile_coord(sm.list[(head + 1u) & 31u] & 0x7FFFFFFFu, a);
    bool tile_nonzero = false, started = false;
#pragma unroll 1
    for (int z1 = 0; z1 < 4; z1++, sc++) {
      if (z1 < 2) issue_slab(sm.ring[(sc + 2) % kRingSlabs], a, c, z1 + 2, lane, q, r);
      else if (have_next) { issue_slab(sm.ring[(sc + 2) % kRingSlabs], a, cn, z1 - 2, lane, q, r); ahead = 2; }
      cp_async_commit();        // one group per slab step, empty or not: "all but the two newest groups" is always the slab consumed now
      cp_async_wait<2>();
      tile_nonzero |= process_slab(sm, sm.ring[sc % kRingSlabs], z1, lane, q, r, z1 == 3, qn, started);
      if (zero_pending) store_zero_part(a, cz, lane, z1);
    }
    zero_pending = false;
    if (tile_nonzero) {
      finish_tile(sm, a, c, lane);
      if (lane == 0) a.tile_zero[tile] = 0;
    } else if (!(entry >> 31)) {
      // empty, but the outputs of the previous build are not known to be zero: write zeros (exact: the filter of zeros is zero)
      zero_pending = true;
      cz = c;
      if (lane == 0) a.tile_zero[tile] = 1;
    }
    // else: empty now and every output (incl. the occupancy words) known to be zero from the previous build: nothing to write.
    // The usual case: < 1 % of the grid is occupied and the occupied set moves little between frames.
    head++;
    if (!sparse) {
      if (pending) {   // collect the tile requested one tile ago, request the next one
        const uint32_t t = __shfl_sync(0xffffffffu, pend, 0);
        pending = false;
        if (t < a.n_tiles) {
          if (lane == 0) sm.list[tail & 31u] = t;
          tail++;
          if (lane == 0) pend = 2u * n_warps + atomicAdd(a.counter, 1u);
          pending = true;
        }
        __syncwarp();
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// occupancy bits
struct OccArgs {
  const uint32_t* src0;                 // level 0 words (generic path)
  uint32_t* occ[VCT_MAX_LEVELS];
  uint32_t* docc[VCT_MAX_LEVELS];
  const uint8_t* occb;                  // fused path: occupancy bytes of levels 3..
  uint32_t occb_off[VCT_MAX_LEVELS];
  int R, levels;
};

__global__ void __launch_bounds__(256)
occ_bits_kernel(const OccArgs a) {   // level 0 from the voxel words (only when the fused kernel did not run)
  const size_t N = (size_t)a.R, n = N * N * N;
  const uint32_t* src = a.src0;
  const int lane = threadIdx.x & 31;
  const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w * 32 < n; w += n_warps) {
    const size_t i = w * 32 + lane;
    const uint32_t any = i < n ? src[i] : 0u;
    const uint32_t bal = __ballot_sync(0xffffffffu, any != 0u);
    if (lane == 0) a.occ[0][w] = bal;
  }
}

// occupancy bit (x,y,z) of a level (flat index (z*N + y)*N + x)
__device__ __forceinline__ uint32_t occ_bit(const uint32_t* __restrict__ occ, int N, int x, int y, int z) {
  const size_t flat = ((size_t)z * N + y) * N + x;
  return (occ[flat >> 5] >> (flat & 31)) & 1u;
}

// Occupancy of a level from the level below: bit = OR of the 8 child bits (= "a voxel of the texel's level-0 support is non-zero").
// R is a power of two (vct_grid_create), so every index below is shifts and masks.
// N >= 32: an output word is 32 texels of one row = two source words in each of four source rows.
__device__ __forceinline__ uint32_t occ_reduce_word(const uint32_t* __restrict__ src, int logN, uint32_t w) {
  const int lw = logN - 5;                                   // log2(words per destination row)
  const uint32_t k = w & ((1u << lw) - 1u), row = w >> lw, y = row & ((1u << logN) - 1u), z = row >> logN;
  uint32_t lo = 0, hi = 0;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const uint32_t* r = src + ((((size_t)(2 * z + (q >> 1)) << (logN + 1)) + (2 * y + (q & 1))) << (lw + 1)) + 2 * k;
    lo |= r[0]; hi |= r[1];
  }
  return occ_pair_or(lo) | (occ_pair_or(hi) << 16);
}
// N < 32: one thread per texel (flat bit order), the warp's ballot is the output word
__device__ __forceinline__ uint32_t occ_reduce_texel(const uint32_t* __restrict__ src, int logN, uint32_t i) {
  const uint32_t m = (1u << logN) - 1u, x = i & m, y = (i >> logN) & m, z = i >> (2 * logN);
  const int ls = logN + 1;                                   // source level: Ns = 2N <= 32 texels per row
  uint32_t any = 0;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const uint32_t flat = ((((2 * z + (q >> 1)) << ls) + (2 * y + (q & 1))) << ls) + 2 * x;   // children (2x, .) and (2x+1, .): same word
    any |= (src[flat >> 5] >> (flat & 31)) & 3u;
  }
  return any;
}

// one level with N >= 64: a word per thread over the whole grid
__global__ void __launch_bounds__(256)
occ_reduce_level_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int logN) {
  const uint32_t n_words = 1u << (3 * logN - 5);
  for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += gridDim.x * blockDim.x) dst[w] = occ_reduce_word(src, logN, w);
}

// the small levels (N <= 32) depend on each other and are tiny: ONE CTA walks them in order (generic path only)
__global__ void __launch_bounds__(1024)
occ_reduce_kernel(const OccArgs a, int first_level, int levels) {
  for (int level = first_level; level < levels; level++) {
    const int N = a.R >> level;
    int logN = 0;
    while ((1 << logN) < N) logN++;
    const uint32_t* __restrict__ src = a.occ[level - 1];
    uint32_t* __restrict__ dst = a.occ[level];
    if (N >= 32) {
      const uint32_t n_words = 1u << (3 * logN - 5);
      for (uint32_t w = threadIdx.x; w < n_words; w += blockDim.x) dst[w] = occ_reduce_word(src, logN, w);
    } else {
      const uint32_t n = 1u << (3 * logN);
      for (uint32_t base = 0; base < n; base += blockDim.x) {   // uniform per warp: every lane takes part in the ballot
        const uint32_t i = base + threadIdx.x;
        const uint32_t any = i < n ? occ_reduce_texel(src, logN, i) : 0u;
        const uint32_t bal = __ballot_sync(0xffffffffu, any != 0u);
        if ((threadIdx.x & 31) == 0 && i < n) dst[i >> 5] = bal;
      }
    }
    __syncthreads();   // the next level reads what this CTA just wrote
  }
}

// 32 occupancy bits of row (y,z) starting at x = 32*k; rows outside the level read as zero
__device__ __forceinline__ uint32_t occ_row_bits(const uint32_t* __restrict__ occ, int N, int y, int z, int k) {
  if ((unsigned)y >= (unsigned)N || (unsigned)z >= (unsigned)N || k < 0 || k * 32 >= N) return 0u;
  const size_t flat = ((size_t)z * N + y) * N + (size_t)k * 32;
  if (N >= 32) return occ[flat >> 5];
  return (occ[flat >> 5] >> (flat & 31)) & ((1u << N) - 1u);
}

// dilation: docc bit (x+1,y+1,z+1) = OR of occ over [x,x+1]x[y,y+1]x[z,z+1].  One thread per output ROW (yy,zz) of a level:
// walks the row's words carrying the top bit of the previous word.
__device__ __forceinline__ void occ_dilate_row(const OccArgs& a, int level, int yy, int zz) {
  const int N = a.R >> level, D = N + 1, wpr = occ_wpr(N);
  const uint32_t* occ = a.occ[level];
  uint32_t* __restrict__ out = a.docc[level] + ((size_t)zz * D + yy) * wpr;
  const int y = yy - 1, z = zz - 1;
  if (N >= 32) {
    const int nw = N >> 5;
    const uint32_t* rows[4];
    bool ok[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int ry = y + (q & 1), rz = z + (q >> 1);
      ok[q] = (unsigned)ry < (unsigned)N && (unsigned)rz < (unsigned)N;
      rows[q] = occ + ((size_t)(ok[q] ? rz : 0) * N + (ok[q] ? ry : 0)) * nw;
    }
    uint32_t carry = 0;
    for (int k = 0; k < wpr; k++) {
      uint32_t r = 0;
      if (k < nw) {
#pragma unroll
        for (int q = 0; q < 4; q++) r |= ok[q] ? rows[q][k] : 0u;
      }
      out[k] = (r << 1) | carry | r;
      carry = r >> 31;
    }
  } else {
    uint32_t r = 0;
#pragma unroll
    for (int dz = 0; dz < 2; dz++)
#pragma unroll
      for (int dy = 0; dy < 2; dy++) r |= occ_row_bits(occ, N, y + dy, z + dz, 0);
    out[0] = (r << 1) | r;
  }
}

// generic path. grid: x = rows of one z-slice, y = z-slice, z = level
__global__ void __launch_bounds__(128)
occ_dilate_kernel(const OccArgs a) {
  const int level = (int)blockIdx.z;
  const int D = (a.R >> level) + 1;
  const int zz = (int)blockIdx.y, yy = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (zz >= D || yy >= D) return;
  occ_dilate_row(a, level, yy, zz);
}

// one dilated word: bits 32 k .. 32 k + 31 of row (yy, zz); reads the word and its left neighbour in four rows
__device__ __forceinline__ void occ_dilate_word(const OccArgs& a, int level, int k, int yy, int zz) {
  const int N = a.R >> level, D = N + 1, wpr = occ_wpr(N);
  const uint32_t* __restrict__ occ = a.occ[level];
  const int y = yy - 1, z = zz - 1;
  uint32_t r = 0, c = 0;
  if (N >= 32) {
    const int nw = N >> 5;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int ry = y + (q & 1), rz = z + (q >> 1);
      if ((unsigned)ry < (unsigned)N && (unsigned)rz < (unsigned)N) {
        const uint32_t* row = occ + ((size_t)rz * N + ry) * nw;
        if (k < nw) r |= row[k];
        if (k >= 1) c |= row[k - 1];
      }
    }
  } else {
#pragma unroll
    for (int q = 0; q < 4; q++) r |= occ_row_bits(occ, N, y + (q & 1), z + (q >> 1), 0);
  }
  a.docc[level][((size_t)zz * D + yy) * wpr + k] = (r << 1) | r | (c >> 31);
}

// Eight dilated words of one column: word k of rows yy0 .. yy0 + 7 of slice zz.  The 36 source words (two words of nine rows of
// two slices; every source row feeds two output rows) are loaded up front: one memory latency per eight outputs.  (One word per
// thread with its eight loads was latency-bound at 3 resident blocks per SM: 12 of the 17 us of the tail kernel at 256^3.)
__device__ __forceinline__ void occ_dilate_rows8(const OccArgs& a, int level, int k, int yy0, int zz) {
  const int N = a.R >> level, D = N + 1, wpr = occ_wpr(N);
  if (N < 32) {
    for (int j = 0; j < 8 && yy0 + j < D; j++) occ_dilate_word(a, level, k, yy0 + j, zz);
    return;
  }
  const uint32_t* __restrict__ occ = a.occ[level];
  const int nw = N >> 5;
  uint32_t r[9], c[9];
  if (yy0 >= 1 && yy0 + 7 < N && zz >= 1 && zz < N && k >= 1 && k < nw) {
    // interior (almost every column): 36 loads at constant strides from one base pointer, no clamping, no masks
    const uint32_t* p = occ + ((size_t)(zz - 1) * N + (yy0 - 1)) * nw + (k - 1);
    const size_t sz = (size_t)N * nw;
    uint2 lo[9], hi[9];
#pragma unroll
    for (int j = 0; j < 9; j++) {
      lo[j] = make_uint2(__ldg(p + (size_t)j * nw), __ldg(p + (size_t)j * nw + 1));
      hi[j] = make_uint2(__ldg(p + sz + (size_t)j * nw), __ldg(p + sz + (size_t)j * nw + 1));
    }
#pragma unroll
    for (int j = 0; j < 9; j++) { c[j] = lo[j].x | hi[j].x; r[j] = lo[j].y | hi[j].y; }
  } else {
  // every load is unconditional (clamped address, result masked): 36 independent loads in flight, no branch between them
  const int kr = min(k, nw - 1), kc = max(k - 1, 0);
  const uint32_t mr = k < nw ? 0xFFFFFFFFu : 0u, mc = k >= 1 ? 0xFFFFFFFFu : 0u;
  uint32_t lr[9][2], lc[9][2];
#pragma unroll
  for (int j = 0; j < 9; j++) {
    const int y = yy0 - 1 + j, yc = min(max(y, 0), N - 1);
#pragma unroll
    for (int dz = 0; dz < 2; dz++) {
      const int z = zz - 1 + dz, zc = min(max(z, 0), N - 1);
      const uint32_t* row = occ + ((size_t)zc * N + yc) * nw;
      lr[j][dz] = __ldg(row + kr);
      lc[j][dz] = __ldg(row + kc);
    }
  }
#pragma unroll
  for (int j = 0; j < 9; j++) {
    const int y = yy0 - 1 + j;
    const uint32_t my = (unsigned)y < (unsigned)N ? 0xFFFFFFFFu : 0u;
    const uint32_t mz0 = zz >= 1 ? 0xFFFFFFFFu : 0u, mz1 = zz < N ? 0xFFFFFFFFu : 0u;
    r[j] = ((lr[j][0] & mz0) | (lr[j][1] & mz1)) & my & mr;
    c[j] = ((lc[j][0] & mz0) | (lc[j][1] & mz1)) & my & mc;
  }
  }
  uint32_t* __restrict__ out = a.docc[level] + ((size_t)zz * D + yy0) * wpr + k;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    if (yy0 + j < D) {
      const uint32_t rr = r[j] | r[j + 1], cc = c[j] | c[j + 1];
      out[(size_t)j * wpr] = (rr << 1) | rr | (cc >> 31);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Fused path, everything that follows the fused kernel in ONE wave of independent work (mip_tail_kernel):
//   blocks [0, n_sb)                 fold one super-block (64^3 voxels, or the whole grid when R = 32): its level-3 texels (linear
//                                    records of the fused kernel) -> levels 4, 5, 6 of the six directions + their occupancy bytes.
//                                    (bits of the occupied texels ORed into words the fused kernel zeroed).  The CTA that finishes last
//                                    builds levels 7..
//   blocks [n_sb, n_sb + word_blocks) dilation of levels 0-2 from their occupancy words, one thread per column of eight output words
//   the rest                          level 3: one WARP per output word, straight from the occupancy bytes -- lane j = texel 32 k + j:
//                                    plain bits (ballot of the byte) and dilated bits (ballot of the OR of the 2x2x2 bytes)
struct TailArgs {
  OccArgs occ;
  const uint32_t* rec3;        // level 3, records of six words
  uint32_t* rec_top;           // levels 5..: records at top_off[l] (scratch of this kernel)
  uint32_t top_off[VCT_MAX_LEVELS];
  uint8_t* occb;               // occupancy bytes of levels 3.. (level 3 written by the fused kernel, 4.. here)
  uint32_t* ticket;            // super-blocks finished
  uint32_t* stream_counter;    // work counter of the streaming kernel: reset here for the next build
  const uint32_t* sb_epoch;    // per 64^3 super-block: the last build whose streaming kernel processed one of its tiles
  uint32_t build;              // this build: a super-block with another number has unchanged level-0 content (all zero, then and now)
  int log_nsb;
  SurfSet surf;
  int n_sb, sb3;               // super-blocks; level-3 texels per super-block side (8, or 4 when R = 32)
  int row_start[4];            // first block (relative to n_sb) of the dilated words of level 0, 1, 2 (and the end)
  int chunks[3];               // blocks per z-slice of those levels: a block = 256 consecutive 8-row columns of ONE slice (no per-thread division by D)
  int word_warps3;             // warps of the level-3 words (plain, then dilated)
};

// plain and dilated occupancy words of one level from its occupancy bytes; `warp` counts through the plain words, then the dilated ones
__device__ __forceinline__ void occ_words_from_bytes(const OccArgs& a, const uint8_t* __restrict__ ob, int level, int warp, int lane) {
  const int N = a.R >> level, D = N + 1, wpr = occ_wpr(N);
  const int n_plain = (int)occ_words(N);
  if (warp < n_plain) {   // plain bits, flat order
    const size_t i = (size_t)warp * 32 + lane;
    const uint32_t v = i < (size_t)N * N * N ? __ldcg(ob + i) : 0u;
    const uint32_t bal = __ballot_sync(0xffffffffu, v != 0u);
    if (lane == 0) a.occ[level][warp] = bal;
    return;
  }
  const int w = warp - n_plain;   // dilated word k of row (yy, zz)
  if (w >= D * D * wpr) return;
  const int k = w % wpr, row = w / wpr, yy = row % D, zz = row / D;
  const int xx = 32 * k + lane;   // bit (xx, yy, zz) = OR over texels [xx-1, xx] x [yy-1, yy] x [zz-1, zz]
  uint32_t v = 0;
  if (xx <= N) {
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int x = xx - 1 + (q & 1), y = yy - 1 + ((q >> 1) & 1), z = zz - 1 + (q >> 2);
      if ((unsigned)x < (unsigned)N && (unsigned)y < (unsigned)N && (unsigned)z < (unsigned)N) v |= __ldcg(ob + ((size_t)z * N + y) * N + x);
    }
  }
  const uint32_t bal = __ballot_sync(0xffffffffu, v != 0u);
  if (lane == 0) a.docc[level][((size_t)zz * D + yy) * wpr + k] = bal;
}
// 32 occupancy bytes (each 0 or 1, 32-byte aligned) -> 32 bits.  Four bytes b0..b3 of a word land on bits 24..27 of word * 0x01020408
// (every partial product hits a distinct bit position: no carries).
__device__ __forceinline__ uint32_t occ_pack32(const uint8_t* __restrict__ p) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(p)), b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t bits = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) bits |= ((w[i] * 0x01020408u) >> 24 & 0xFu) << (4 * i);
  return bits;
}
// The same words with ONE THREAD per word (levels of at least 32 texels per row: a word is a piece of one row).  A warp per word
// with a byte load per lane was 250 warp instructions per word, 17 % of the tail kernel at 1024^3.
__device__ __forceinline__ void occ_word_from_bytes_thread(const OccArgs& a, const uint8_t* __restrict__ ob, int level, int w) {
  const int N = a.R >> level, D = N + 1, wpr = occ_wpr(N), nw = N >> 5;
  const int n_plain = (int)occ_words(N);
  if (w < n_plain) { a.occ[level][w] = occ_pack32(ob + (size_t)w * 32); return; }
  w -= n_plain;   // dilated word k of row (yy, zz): bit j = OR over texels [32k+j-1, 32k+j] x [yy-1, yy] x [zz-1, zz]
  if (w >= D * D * wpr) return;
  const int k = w % wpr, row = w / wpr, yy = row % D, zz = row / D;
  uint32_t r = 0, carry = 0;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int y = yy - 1 + (q & 1), z = zz - 1 + (q >> 1);
    if ((unsigned)y < (unsigned)N && (unsigned)z < (unsigned)N) {
      const uint8_t* rowp = ob + ((size_t)z * N + y) * N;
      if (k < nw) r |= occ_pack32(rowp + 32 * k);
      if (k >= 1) carry |= __ldg(rowp + 32 * k - 1);
    }
  }
  a.docc[level][((size_t)zz * D + yy) * wpr + k] = (r << 1) | r | (carry & 1u);
}
// an occupied texel of a level >= 4: its plain bit and the eight dilated bits that cover it (the words were zeroed by the fused kernel)
__device__ __forceinline__ void occ_set_texel(const OccArgs& a, int level, int x, int y, int z) {
  const int N = a.R >> level, D = N + 1, wpr = occ_wpr(N);
  const size_t flat = ((size_t)z * N + y) * N + x;
  atomicOr(a.occ[level] + (flat >> 5), 1u << (flat & 31));
#pragma unroll
  for (int q = 0; q < 8; q++) {
    const int X = x + (q & 1), Y = y + ((q >> 1) & 1), Z = z + (q >> 2);   // dilated bit (X,Y,Z) covers texels X-1..X, Y-1..Y, Z-1..Z
    atomicOr(a.docc[level] + ((size_t)Z * D + Y) * wpr + (X >> 5), 1u << (X & 31));
  }
}
__host__ __device__ __forceinline__ int occ_word_warps(int N) { return (int)occ_words(N) + (N + 1) * (N + 1) * occ_wpr(N); }

struct TailSmem {
  uint32_t f[6][8][8][8];      // the super-block's level 3 (12 KB), then reused level by level
  uint32_t g[6][4][4][4];
  uint32_t h[6][2][2][2];
  uint8_t o3[8][8][8], o4[4][4][4], o5[2][2][2];
  uint32_t last;
};

// one level inside the super-block: src = n^3 texels per direction in shared memory -> (n/2)^3, stored to shared memory, the array
// and (records) to rec_out when given; occupancy bytes alongside
template <int NS>
__device__ __forceinline__ void fold_level(const uint32_t (&src)[6][NS][NS][NS], uint32_t (*dst)[NS / 2][NS / 2][NS / 2], const uint8_t (&osrc)[NS][NS][NS],
                                           uint8_t (*odst)[NS / 2][NS / 2], const TailArgs& a, int level, int ox, int oy, int oz, uint32_t* rec_out) {
  constexpr int ND = NS / 2;
  const int N = a.occ.R >> level;
  for (int u = (int)threadIdx.x; u < ND * ND * ND * 6; u += (int)blockDim.x) {
    const int tex = u % (ND * ND * ND), d = u / (ND * ND * ND), x = tex % ND, y = (tex / ND) % ND, z = tex / (ND * ND);
    uint32_t w[8];
    uint32_t ob = 0;
#pragma unroll
    for (int dz = 0; dz < 2; dz++)
#pragma unroll
      for (int dy = 0; dy < 2; dy++)
#pragma unroll
        for (int dx = 0; dx < 2; dx++) {
          w[child_id(dx, dy, dz)] = src[d][2 * z + dz][2 * y + dy][2 * x + dx];
          ob |= osrc[2 * z + dz][2 * y + dy][2 * x + dx];
        }
    const uint32_t o = filter1_fp32(w, d);
    if (dst) dst[d][z][y][x] = o;
    surf_write(a.surf, d, level, o, (ox + x) * 4, oy + y, oz + z);
    const size_t gi = ((size_t)(oz + z) * N + (oy + y)) * N + (ox + x);
    if (rec_out) __stcg(rec_out + gi * 6 + d, o);
    if (d == 0) {
      if (odst) odst[z][y][x] = (uint8_t)(ob != 0u);
      __stcg(a.occb + a.occ.occb_off[level] + gi, (uint8_t)(ob != 0u));
      if (ob) occ_set_texel(a.occ, level, ox + x, oy + y, oz + z);
    }
  }
}

// levels `first`.. from the records of level first - 1, by ONE CTA (a few hundred texels)
__device__ void tail_top(const TailArgs& a, int first) {
  for (int l = first; l < a.occ.levels; l++) {
    const int N = a.occ.R >> l, Ns = N * 2;
    const uint32_t* src = a.rec_top + a.top_off[l - 1];
    uint32_t* dst = a.rec_top + a.top_off[l];
    uint8_t* ob = a.occb + a.occ.occb_off[l];
    const uint8_t* ob_src = a.occb + a.occ.occb_off[l - 1];
    const int n = N * N * N * 6;
    for (int u = (int)threadIdx.x; u < n; u += (int)blockDim.x) {
      const int d = u % 6, tex = u / 6, x = tex % N, y = (tex / N) % N, z = tex / (N * N);
      uint32_t w[8];
      uint32_t ob_any = 0;
#pragma unroll
      for (int dz = 0; dz < 2; dz++)
#pragma unroll
        for (int dy = 0; dy < 2; dy++)
#pragma unroll
          for (int dx = 0; dx < 2; dx++) {
            const size_t si = ((size_t)(2 * z + dz) * Ns + (2 * y + dy)) * Ns + (2 * x + dx);
            w[child_id(dx, dy, dz)] = __ldcg(src + si * 6 + d);
            ob_any |= __ldcg(ob_src + si);
          }
      const uint32_t o = filter1_fp32(w, d);
      __stcg(dst + (size_t)tex * 6 + d, o);
      surf_write(a.surf, d, l, o, x * 4, y, z);
      if (d == 0) {
        __stcg(ob + tex, (uint8_t)(ob_any != 0u));
        if (ob_any) occ_set_texel(a.occ, l, x, y, z);
      }
    }
    __syncthreads();   // the next level reads what this CTA just wrote (global memory, L2 loads)
  }
}

// the super-block's level 3 into shared memory: 12 (SB3 = 8) or 2 words per thread, every load issued before the first one is used (one
// L2 latency on the critical path of the launch instead of twelve); SB3 is a compile-time constant: the index arithmetic is shifts
template <int SB3>
__device__ __forceinline__ void load_super_block(TailSmem& sm, const TailArgs& a, const uint8_t* __restrict__ ob3, int sx, int sy, int sz, int N3, int t) {
  constexpr int kWords = SB3 * SB3 * SB3 * 6, kIter = (kWords + 255) / 256;
  uint32_t v[kIter];
  uint8_t ob[kIter];
#pragma unroll
  for (int i = 0; i < kIter; i++) {
    const int u = t + 256 * i;
    v[i] = 0u; ob[i] = 0;
    if (u < kWords) {
      const int d = u % 6, tex = u / 6, x = tex % SB3, y = (tex / SB3) % SB3, z = tex / (SB3 * SB3);
      const size_t gi = ((size_t)(sz * SB3 + z) * N3 + (sy * SB3 + y)) * N3 + (sx * SB3 + x);
      v[i] = __ldcg(a.rec3 + gi * 6 + d);
      if (d == 0) ob[i] = __ldcg(ob3 + gi);
    }
  }
#pragma unroll
  for (int i = 0; i < kIter; i++) {
    const int u = t + 256 * i;
    if (u < kWords) {
      const int d = u % 6, tex = u / 6, x = tex % SB3, y = (tex / SB3) % SB3, z = tex / (SB3 * SB3);
      sm.f[d][z][y][x] = v[i];
      if (d == 0) sm.o3[z][y][x] = ob[i];
    }
  }
}

// does the box of level-`level` texels [x0,x1] x [y0,y1] x [z0,z1] (clamped to the level) touch a super-block that the streaming kernel
// processed in this build?  If not, everything derived from those texels is unchanged since the last build (and already in memory).
__device__ __forceinline__ bool region_active(const TailArgs& a, int level, int x0, int x1, int y0, int y1, int z0, int z1) {
  const int N = a.occ.R >> level, sh = 6 - level;   // a super-block is 64 >> level texels wide
  const int nsb = 1 << a.log_nsb;
  const int sx0 = min(max(x0, 0) >> sh, nsb - 1), sx1 = min(min(x1, N - 1) >> sh, nsb - 1);
  const int sy0 = min(max(y0, 0) >> sh, nsb - 1), sy1 = min(min(y1, N - 1) >> sh, nsb - 1);
  const int sz0 = min(max(z0, 0) >> sh, nsb - 1), sz1 = min(min(z1, N - 1) >> sh, nsb - 1);
  for (int z = sz0; z <= sz1; z++)
    for (int y = sy0; y <= sy1; y++)
      for (int x = sx0; x <= sx1; x++)
        if (__ldg(a.sb_epoch + (((z << a.log_nsb) + y) << a.log_nsb) + x) == a.build) return true;
  return false;
}

__global__ void __launch_bounds__(256)
mip_tail_kernel(const TailArgs a) {
  __shared__ TailSmem sm;
  const int b = (int)blockIdx.x, t = (int)threadIdx.x;
  if (b == 0 && t == 0) *a.stream_counter = 0u;
  if (b < a.n_sb) {
    const int R = a.occ.R, N3 = R >> 3, sb3 = a.sb3, nsb = N3 / sb3;
    const int sx = b % nsb, sy = (b / nsb) % nsb, sz = b / (nsb * nsb);
    const uint8_t* ob3 = a.occb + a.occ.occb_off[3];
    if (a.sb_epoch[b] != a.build) {
      // untouched super-block: its colours of levels 4.. are still in memory; only the occupancy bits of those levels, which the
      // streaming kernel zeroes every build, are set again from the occupancy bytes of the last fold
      const int top = min(a.occ.levels, sb3 == 8 ? 7 : 6);
      for (int l = 4, n = sb3 / 2; l < top; l++, n >>= 1) {
        const int N = R >> l;
        for (int u = t; u < n * n * n; u += 256) {
          const int x = sx * n + u % n, y = sy * n + (u / n) % n, z = sz * n + u / (n * n);
          if (__ldcg(a.occb + a.occ.occb_off[l] + ((size_t)z * N + y) * N + x)) occ_set_texel(a.occ, l, x, y, z);
        }
      }
      if (a.occ.levels <= (sb3 == 8 ? 7 : 6)) return;
    } else {
    if (sb3 == 8) load_super_block<8>(sm, a, ob3, sx, sy, sz, N3, t);
    else load_super_block<4>(sm, a, ob3, sx, sy, sz, N3, t);
    __syncthreads();
    // a record copy of the last folded level feeds the top of the chain
    if (sb3 == 8) {
      fold_level<8>(sm.f, sm.g, sm.o3, sm.o4, a, 4, sx * 4, sy * 4, sz * 4, nullptr);
      __syncthreads();
      fold_level<4>(sm.g, sm.h, sm.o4, sm.o5, a, 5, sx * 2, sy * 2, sz * 2, nullptr);
      __syncthreads();
      if (a.occ.levels > 6) fold_level<2>(sm.h, nullptr, sm.o5, nullptr, a, 6, sx, sy, sz, a.rec_top + a.top_off[6]);
    } else {   // R = 32: the super-block is the grid, 4^3 level-3 texels -> levels 4 and 5 (sm.g / sm.o4 hold level 3 here)
      for (int u = t; u < 64 * 6; u += 256) {
        const int d = u / 64, tex = u % 64;
        sm.g[d][tex >> 4][(tex >> 2) & 3][tex & 3] = sm.f[d][tex >> 4][(tex >> 2) & 3][tex & 3];
        if (d == 0) sm.o4[tex >> 4][(tex >> 2) & 3][tex & 3] = sm.o3[tex >> 4][(tex >> 2) & 3][tex & 3];
      }
      __syncthreads();
      fold_level<4>(sm.g, sm.h, sm.o4, sm.o5, a, 4, 0, 0, 0, nullptr);
      __syncthreads();
      fold_level<2>(sm.h, nullptr, sm.o5, nullptr, a, 5, 0, 0, 0, nullptr);
    }
    }
    // ---- the last super-block to finish builds the levels above (none for the usual 7 levels: no fence, no ticket) ----
    if (a.occ.levels <= (sb3 == 8 ? 7 : 6)) return;
    __threadfence();
    __syncthreads();
    if (t == 0) {
      const uint32_t done = atomicAdd(a.ticket, 1u) + 1u;
      sm.last = done == (uint32_t)a.n_sb ? 1u : 0u;
      if (sm.last) { *a.ticket = 0u; __threadfence(); }
    }
    __syncthreads();
    if (sm.last) tail_top(a, sb3 == 8 ? 7 : 6);
    return;
  }
  const int lb0 = b - a.n_sb;
  if (lb0 < a.row_start[3]) {
    int level = 0;
    while (level < 2 && lb0 >= a.row_start[level + 1]) level++;
    const int D = (a.occ.R >> level) + 1, wpr = occ_wpr(a.occ.R >> level);
    const int lb = lb0 - a.row_start[level], zz = lb / a.chunks[level];        // uniform per block
    const int w = (lb - zz * a.chunks[level]) * 256 + t;                        // column of slice zz: word k of row group g
    if (w < ((D + 7) / 8) * wpr) {
      const int g = (int)(((float)w + 0.5f) * (1.0f / (float)wpr));            // exact: w < 2^16, wpr <= 33
      const int k = w - g * wpr;
      // the words depend on texels [32k-1, 32k+31] x [8g-1, 8g+7] x [zz-1, zz]: unchanged (and already written) if no super-block there was processed
      if (region_active(a, level, 32 * k - 1, 32 * k + 31, 8 * g - 1, 8 * g + 7, zz - 1, zz)) occ_dilate_rows8(a.occ, level, k, 8 * g, zz);
    }
    return;
  }
  const int N3 = a.occ.R >> 3;
  if (N3 >= 32) {   // one thread per word; a word is a piece of one row: skipped when no super-block under it was processed
    const int w = (lb0 - a.row_start[3]) * 256 + t;
    if (w < a.word_warps3) {
      const int n_plain = (int)occ_words(N3), nw = N3 >> 5;
      bool act;
      if (w < n_plain) {
        const int k = w % nw, row = w / nw, y = row % N3, z = row / N3;
        act = region_active(a, 3, 32 * k, 32 * k + 31, y, y, z, z);
      } else {
        const int v = w - n_plain, wpr = occ_wpr(N3), D = N3 + 1, k = v % wpr, row = v / wpr, yy = row % D, zz = row / D;
        act = region_active(a, 3, 32 * k - 1, 32 * k + 31, yy - 1, yy, zz - 1, zz);
      }
      if (act) occ_word_from_bytes_thread(a.occ, a.occb + a.occ.occb_off[3], 3, w);
    }
    return;
  }
  const int warp = (lb0 - a.row_start[3]) * 8 + (t >> 5);   // small levels: one warp per word (a word spans rows)
  if (warp < a.word_warps3) occ_words_from_bytes(a.occ, a.occb + a.occ.occb_off[3], 3, warp, t & 31);
}

// ---------------------------------------------------------------------------------------------
// RGBA16F storage variant (vct_grid_create_ex; BASELINE.json config 5): one level per launch, one thread per (destination texel, direction),
// the fp32 recipe of mipmap.comp on exactly converted halves, clamp to [0,1] (the unorm store of the reference), round to half.  A plain
// kernel: the variant is there for coverage and parity, the RGBA8 path above is the one measured against the roofline.
__device__ __forceinline__ void unpack_half4(unsigned long long c, float (&o)[4]) {
#pragma unroll
  for (int k = 0; k < 4; k++) o[k] = __half2float(__ushort_as_half((unsigned short)(c >> (16 * k))));
}
__global__ void mip_f16_kernel(const unsigned long long* const* __restrict__ table, int level_src, int Ns, int Nd, unsigned long long* const* __restrict__ dst_table_rw) {
  const size_t n = (size_t)Nd * Nd * Nd * 6;
  for (size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x; u < n; u += (size_t)gridDim.x * blockDim.x) {
    const int d = (int)(u % 6);
    const size_t tex = u / 6;
    const int x = (int)(tex % Nd), y = (int)((tex / Nd) % Nd), z = (int)(tex / ((size_t)Nd * Nd));
    const unsigned long long* src = table[level_src * 6 + d];
    float c[8][4];
#pragma unroll
    for (int dz = 0; dz < 2; dz++)
#pragma unroll
      for (int dy = 0; dy < 2; dy++)
#pragma unroll
        for (int dx = 0; dx < 2; dx++) unpack_half4(src[((size_t)(2 * z + dz) * Ns + (2 * y + dy)) * Ns + (2 * x + dx)], c[child_id(dx, dy, dz)]);
    const uint32_t word = mip_pair_word(d >> 1);
    const uint32_t fronts = (d & 1) ? word >> 16 : word & 0xFFFFu, backs = (d & 1) ? word & 0xFFFFu : word >> 16;
    unsigned long long out = 0ull;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      float sum = 0.f;
#pragma unroll
      for (int p = 0; p < 4; p++) {
        const int f = (int)((fronts >> (4 * p)) & 7u), b = (int)((backs >> (4 * p)) & 7u);
        // dynamic child index on a register array would go to local memory: select
        float fk = 0.f, fa = 0.f, bk = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) { if (i == f) { fk = c[i][k]; fa = c[i][3]; } if (i == b) bk = c[i][k]; }
        const float v = __fadd_rn(fk, __fmul_rn(__fadd_rn(1.0f, -fa), bk));
        sum = p == 0 ? v : __fadd_rn(sum, v);
      }
      const float r = fminf(fmaxf(__fdiv_rn(sum, 4.0f), 0.0f), 1.0f);
      out |= (unsigned long long)__half_as_ushort(__float2half_rn(r)) << (16 * k);
    }
    dst_table_rw[(level_src + 1) * 6 + d][tex] = out;
  }
}
__global__ void __launch_bounds__(256)
occ_bits_f16_kernel(const unsigned long long* __restrict__ src, uint32_t* __restrict__ occ0, size_t n) {
  const int lane = threadIdx.x & 31;
  const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w * 32 < n; w += n_warps) {
    const size_t i = w * 32 + lane;
    const unsigned long long any = i < n ? src[i] : 0ull;
    const uint32_t bal = __ballot_sync(0xffffffffu, any != 0ull);
    if (lane == 0) occ0[w] = bal;
  }
}

bool mip_fused_applies(int R, int levels) { return R >= 32 && levels >= 6; }

int launch_mipmap(vct_device* dev, vct_grid* g) {
  cudaStream_t s = dev->stream;
  const int R = g->R;
  OccArgs oa;
  memset(&oa, 0, sizeof oa);
  oa.R = R; oa.levels = g->levels; oa.src0 = g->base; oa.occb = g->occb;
  for (int l = 0; l < VCT_MAX_LEVELS; l++) { oa.occ[l] = g->occ[l]; oa.docc[l] = g->docc[l]; oa.occb_off[l] = g->occb_off[l]; }

  if (g->fmt == VCT_GRID_RGBA16F) {
    for (int l = 0; l + 1 < g->levels; l++) {
      const int Ns = max(R >> l, 1), Nd = R >> (l + 1);
      if (Nd < 1) break;
      const size_t n = (size_t)Nd * Nd * Nd * 6;
      mip_f16_kernel<<<grid_for(n), 256, 0, s>>>(g->f16_table, l, Ns, Nd, const_cast<unsigned long long* const*>(reinterpret_cast<const unsigned long long* const*>(g->f16_table)));
    }
    occ_bits_f16_kernel<<<grid_for((size_t)R * R * R, 256, 148 * 8), 256, 0, s>>>(reinterpret_cast<const unsigned long long*>(g->base), g->occ[0], (size_t)R * R * R);
  } else if (mip_fused_applies(R, g->levels)) {
    StreamArgs fa;
    memset(&fa, 0, sizeof fa);
    fa.base = g->base; fa.R = R; fa.levels = g->levels; fa.surf = g->surf;
    fa.n_tiles = (uint32_t)((R / WX) * (R / WY) * (R / WZ));
    while ((WX << fa.log_tx) < R) fa.log_tx++;
    while ((WY << fa.log_ty) < R) fa.log_ty++;
    fa.occ0 = g->occ[0]; fa.occ1 = (uint16_t*)g->occ[1]; fa.occ2 = (uint8_t*)g->occ[2];
    fa.rec3 = g->rec3; fa.occb3 = g->occb + g->occb_off[3];
    fa.zero0 = g->occ_hi; fa.zero0_n = g->occ_hi_words; fa.zero1 = g->docc_all + g->docc_hi_off; fa.zero1_n = g->docc_hi_words;
    fa.tile_zero = g->tile_zero;
    fa.counter = g->mip_counters + 1;
    int log_nsb = 0;
    while ((64 << log_nsb) < R) log_nsb++;
    g->mip_build++;
    fa.sb_epoch = g->sb_epoch; fa.build = g->mip_build; fa.log_nsb = log_nsb;
    fa.touched = (g->flags_valid && !g->external) ? g->tile_touched : nullptr;
    if (g->peer_touched) fa.touched = g->peer_touched;   // multi-GPU frame: flags kept by the ranks that stored the voxels (peer.cu)
    if (dev->debug_mip_dense) {   // measurement switch (vct_debug_set): the dense build, every tile read and written
      fa.touched = nullptr;
      fa.dense = 1;
    }
    const int smem = (int)(sizeof(WarpSmem) * kStreamWarps);
    if (!dev->mip_attr_set) {
      VCT_CUDA(cudaFuncSetAttribute(mip_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      dev->mip_attr_set = true;
    }
    // persistent: 2 CTAs of 6 warps per SM (17 KB of shared memory per warp), fewer when there are fewer tiles than warps
    const int want = (int)((fa.n_tiles + kStreamWarps - 1) / kStreamWarps);
    const int ctas = min(want, dev->prop.multiProcessorCount * 2);
    mip_stream_kernel<<<ctas, kStreamThreads, smem, s>>>(fa);
    TailArgs ta;
    memset(&ta, 0, sizeof ta);
    ta.occ = oa; ta.rec3 = g->rec3; ta.rec_top = g->rec_top; ta.occb = g->occb; ta.ticket = g->mip_counters; ta.stream_counter = g->mip_counters + 1; ta.surf = g->surf;
    ta.sb_epoch = g->sb_epoch; ta.build = g->mip_build; ta.log_nsb = log_nsb;
    for (int l = 0; l < VCT_MAX_LEVELS; l++) ta.top_off[l] = g->top_off[l];
    ta.sb3 = R >= 64 ? 8 : 4;
    ta.n_sb = R >= 64 ? (R / 64) * (R / 64) * (R / 64) : 1;
    int blocks = 0;
    for (int l = 0; l < 3; l++) {
      const int D = (R >> l) + 1;
      ta.row_start[l] = blocks;
      ta.chunks[l] = (((D + 7) / 8) * occ_wpr(R >> l) + 255) / 256;
      blocks += D * ta.chunks[l];
    }
    ta.row_start[3] = blocks;
    ta.word_warps3 = occ_word_warps(R >> 3);
    // level-3 words: one thread per word when a row has at least 32 texels, else one warp per word
    const int blocks3 = (R >> 3) >= 32 ? (ta.word_warps3 + 255) / 256 : (ta.word_warps3 + 7) / 8;
    mip_tail_kernel<<<ta.n_sb + blocks + blocks3, 256, 0, s>>>(ta);
    VCT_CUDA(cudaGetLastError());
    return VCT_OK;
  }

  // ---- generic path (small grids / short chains): one launch per level ----
  for (int l = 0; l + 1 < g->levels && g->fmt == VCT_GRID_RGBA8; l++) {
    const int Ns = max(R >> l, 1), Nd = R >> (l + 1);
    if (Nd < 1) break;
    const size_t n = (size_t)Nd * Nd * Nd * 6;
    mip_generic_kernel<<<grid_for(n), 256, 0, s>>>(l == 0 ? g->base : nullptr, l == 0 ? 0 : g->surf.s[l], l == 0 ? 0 : g->surf.pitch[l], g->surf.s[l + 1],
                                                  g->surf.pitch[l + 1], Ns, Nd);
  }
  if (g->fmt == VCT_GRID_RGBA8) occ_bits_kernel<<<grid_for((size_t)R * R * R, 256, 148 * 8), 256, 0, s>>>(oa);
  int small_first = 1;
  for (; small_first < g->levels && (R >> small_first) >= 64; small_first++) {
    int logN = 0;
    while ((1 << logN) < (R >> small_first)) logN++;
    const size_t n_words = (size_t)1 << (3 * logN - 5);
    occ_reduce_level_kernel<<<grid_for(n_words), 256, 0, s>>>(g->occ[small_first - 1], g->occ[small_first], logN);
  }
  if (small_first < g->levels) occ_reduce_kernel<<<1, 1024, 0, s>>>(oa, small_first, g->levels);
  occ_dilate_kernel<<<dim3((R + 1 + 127) / 128, R + 1, g->levels), 128, 0, s>>>(oa);
  VCT_CUDA(cudaGetLastError());
  return VCT_OK;
}

}  // namespace vct
