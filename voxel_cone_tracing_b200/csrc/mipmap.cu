// mipmap.cu -- six-direction anisotropic mip chain of the voxel grid.
//
// Replaces Renderer::filter() (src/renderer.cpp:283-314) + shader/mipmap.comp.  The reference dispatches one compute
// pass per level, each re-reading six full textures (and launching 8x more threads than texels).  Here the whole chain
// -- any number of levels -- is built by TWO launches:
//   mip_fused_kernel   persistent CTAs stream 32x16x8 tiles of level 0 through a 3-stage TMA ring
//                      (cp.async.bulk.tensor.3d + mbarrier; level 0 is read ONCE, not 6x) -> levels 1, 2, 3 of all six
//                      directions and the occupancy bits of levels 0-2.  Pure streaming: no atomics, no fences, no
//                      dependence between CTAs.  Per round a CTA first compacts the tiles it has to read: with the
//                      voxelizer's tile flags, untouched tiles whose outputs are already zero are neither read nor written.
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
#include <cuda.h>

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
// the fused kernel.  Tile = 32 x 16 x 8 level-0 texels (16 KB, rows of 128 bytes); 8 tiles (2 in y, 4 in z) = one 32^3 block.
constexpr int TX = 32, TY = 16, TZ = 8;
constexpr int kStages = 3;
constexpr int kListMax = 512;                 // tiles of one CTA examined per round
constexpr uint32_t kTileBytes = TX * TY * TZ * 4;
constexpr int kFusedThreads = 256;

struct FusedArgs {
  int R, levels;
  int n_tiles, log_tx, log_ty;     // tiles per axis: R/32, R/16 (and R/8), as shifts
  uint32_t* occ0; uint16_t* occ1; uint8_t* occ2;
  uint8_t* tile_zero;        // per tile: 1 = every output of this tile is known to be zero
  const uint8_t* touched;    // per tile: the voxelizer wrote into it since the last clear (nullptr: unknown, every tile is read)
  int dense;                 // measurement switch: every tile is read AND written (nothing is taken as known)
  uint32_t *zero0, *zero1;   // plain / dilated occupancy words of levels 4..: zeroed here, the tail kernel ORs the occupied texels in
  uint32_t zero0_n, zero1_n;
  uint32_t* rec3;            // [N3^3][6]: level 3 once more, linear, for the tail kernel
  uint8_t* occb3;            // [N3^3]: level-3 occupancy bytes
  SurfSet surf;
};

struct FusedSmem {
  uint32_t s0[kStages][TZ][TY][TX];            // 3 x 16 KB, TMA destinations (128-byte aligned)
  uint32_t s1[6][TZ / 2][TY / 2][TX / 2];      // 12 KB, one plane per direction
  uint32_t s2[6][TZ / 4][TY / 4][TX / 4];      // 1.5 KB
  uint32_t s3[6][TY / 8][TX / 8];              // 48 words
  uint32_t occ_rows[TZ][TY];                   // level-0 occupancy word of every row of the tile
  uint32_t list[kListMax];                     // the tiles this CTA has to read (bit 31: tile_zero[] of the tile)
  uint32_t list_n;
  unsigned long long full[kStages];            // mbarriers: "tile landed"
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one thread: arm the barrier with the tile size and start the 3-D tensor copy global -> shared
__device__ __forceinline__ void tma_load_tile(const CUtensorMap* tmap, void* dst, unsigned long long* bar, int x, int y, int z) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(kTileBytes) : "memory");
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
               "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
               : "memory");
}

// R is a power of two (vct_grid_create), so are the tile counts: shifts instead of integer divisions
__device__ __forceinline__ void tile_coords(uint32_t tile, const FusedArgs& a, int& bx, int& by, int& bz) {
  bx = (int)(tile & ((1u << a.log_tx) - 1u));
  by = (int)((tile >> a.log_tx) & ((1u << a.log_ty) - 1u));
  bz = (int)(tile >> (a.log_tx + a.log_ty));
}
// Per-thread store slots, the same for every tile (only the tile origin changes): computed once per CTA, not per tile.
//   level 1: 6 directions x 4 z x 8 y x 4 quads = 768 16-byte stores, three per thread; a warp = one z-slice of one direction
//            (8 rows of 64 bytes)
//   level 2: 6 x 2 x 4 x 2 = 96 stores (threads 64..159), level 3: 6 x 1 x 2 x 1 = 12 stores (threads 0..11)
struct StoreSlots {
  int l1_xq, l1_y, l1_z[3], l1_d[3];
  int l2_xq, l2_y, l2_z, l2_d;
};

// levels 1-3 of an all-zero tile: zero stores
__device__ __forceinline__ void store_tile_levels(const FusedSmem* sm, const FusedArgs& a, const StoreSlots& ss, int x0, int y0, int z0, int t, bool level1, bool level2, bool level3) {
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  if (level1) {
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const int d = ss.l1_d[i], z = ss.l1_z[i];
      const uint4 v = sm ? *reinterpret_cast<const uint4*>(&sm->s1[d][z][ss.l1_y][4 * ss.l1_xq]) : zero;
      surf_write(a.surf, d, 1, v, (x0 / 2 + 4 * ss.l1_xq) * 4, y0 / 2 + ss.l1_y, z0 / 2 + z);
    }
  }
  if (level2 && t >= 64 && t < 160) {
    const uint4 v = sm ? *reinterpret_cast<const uint4*>(&sm->s2[ss.l2_d][ss.l2_z][ss.l2_y][4 * ss.l2_xq]) : zero;
    surf_write(a.surf, ss.l2_d, 2, v, (x0 / 4 + 4 * ss.l2_xq) * 4, y0 / 4 + ss.l2_y, z0 / 4 + ss.l2_z);
  }
  if (level3) {
    const int N3 = a.R >> 3;
    if (t < 12) {
      const int y = t & 1, d = t >> 1;
      const uint4 v = sm ? *reinterpret_cast<const uint4*>(&sm->s3[d][y][0]) : zero;
      surf_write(a.surf, d, 3, v, (x0 / 8) * 4, y0 / 8 + y, z0 / 8);
    }
    if (t >= 32 && t < 80) {   // the linear copy the block stage reads, and the level-3 occupancy bytes
      const int u = t - 32, d = u % 6, tex = u / 6, x = tex & 3, y = tex >> 2;
      __stcg(a.rec3 + (((size_t)(z0 / 8) * N3 + (y0 / 8 + y)) * N3 + (x0 / 8 + x)) * 6 + d, sm ? sm->s3[d][y][x] : 0u);
    }
  }
}

// reduces one staged tile; every thread of the CTA calls it (contains barriers)
__device__ __forceinline__ void process_tile(FusedSmem& sm, const uint32_t (&s0)[TZ][TY][TX], const FusedArgs& a, const StoreSlots& ss, uint32_t tile, uint32_t known_zero,
                                             int bx, int by, int bz) {
  const int R = a.R;
  const int x0 = bx * TX, y0 = by * TY, z0 = bz * TZ;
  const int t = threadIdx.x;
  const int N1 = R >> 1, N2 = R >> 2, N3 = R >> 3;

  // ---- occupancy bits of level 0 and "is the tile empty": 128 rows of 128 B, four 16-byte shared loads per thread ----
  uint32_t any0 = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int row = (t >> 3) + 32 * k, quad = t & 7, y = row & 15, z = row >> 4;
    const uint4 v = *reinterpret_cast<const uint4*>(&s0[z][y][4 * quad]);
    any0 |= (v.x | v.y) | (v.z | v.w);
    uint32_t bits = ((v.x != 0u) | (v.y != 0u) << 1 | (v.z != 0u) << 2 | (v.w != 0u) << 3) << (4 * quad);
    bits |= __shfl_xor_sync(0xffffffffu, bits, 1);
    bits |= __shfl_xor_sync(0xffffffffu, bits, 2);
    bits |= __shfl_xor_sync(0xffffffffu, bits, 4);
    if (quad == 0) sm.occ_rows[z][y] = bits;
  }
  const int tile_nonzero = __syncthreads_or((int)(any0 != 0u));
  // empty now and every output (incl. the occupancy words) known to be zero from the previous build: nothing to write.
  // The usual case: < 1 % of the grid is occupied and the occupied set moves little between frames.
  if (!tile_nonzero && known_zero) return;
  if (t < TZ * TY) a.occ0[((size_t)(z0 + (t >> 4)) * R + (y0 + (t & 15))) * (R / 32) + bx] = sm.occ_rows[t >> 4][t & 15];
  // occupancy of a texel of level >= 1 = "a voxel of its level-0 support is non-zero" (a superset of "the texel is non-zero": a
  // filtered value can round to zero), so that the bits of a level are the OR of the 8 child bits -- the tracer relies on that.
  if (t >= 128 && t < 160) {   // level 1: 32 rows of 16 bits
    const int r = t - 128, y = r & 7, z = r >> 3;
    const uint32_t w = (sm.occ_rows[2 * z][2 * y] | sm.occ_rows[2 * z][2 * y + 1]) | (sm.occ_rows[2 * z + 1][2 * y] | sm.occ_rows[2 * z + 1][2 * y + 1]);
    a.occ1[((size_t)(z0 / 2 + z) * N1 + (y0 / 2 + y)) * (N1 / 16) + bx] = (uint16_t)occ_pair_or(w);
  } else if (t >= 160 && t < 168) {   // level 2: 8 rows of 8 bits
    const int r = t - 160, y = r & 3, z = r >> 2;
    uint32_t w = 0;
#pragma unroll
    for (int q = 0; q < 16; q++) w |= sm.occ_rows[4 * z + (q >> 2)][4 * y + (q & 3)];
    a.occ2[((size_t)(z0 / 4 + z) * N2 + (y0 / 4 + y)) * (N2 / 8) + bx] = (uint8_t)occ_pair_or(occ_pair_or(w));
  } else if (t >= 192 && t < 194) {   // level 3: 4 x 2 texels, one occupancy BYTE each (sub-byte bit rows would be shared between tiles)
    const int y3 = t - 192;
    uint32_t w = 0;
#pragma unroll
    for (int q = 0; q < 64; q++) w |= sm.occ_rows[q >> 3][8 * y3 + (q & 7)];
#pragma unroll
    for (int x = 0; x < 4; x++)
      __stcg(a.occb3 + ((size_t)(z0 / 8) * N3 + (y0 / 8 + y3)) * N3 + (x0 / 8 + x), (uint8_t)(((w >> (8 * x)) & 0xFFu) != 0u));
  }

  if (!tile_nonzero) {
    // empty tile: every output of this tile is zero (exact: the filter of zeros is zero)
    if (t == 0) a.tile_zero[tile] = 1;
    store_tile_levels(nullptr, a, ss, x0, y0, z0, t, true, true, true);
    return;
  }

  if (t == 0) a.tile_zero[tile] = 0;
  // ---- level 1: two texels per thread (z and z + 2), six directions each ----
  {
    const int x = t & 15, y = (t >> 4) & 7;
#pragma unroll 1
    for (int h = 0; h < 2; h++) {
      const int z = (t >> 7) + 2 * h;
      uint32_t w[8];
      uint32_t any = 0;
#pragma unroll
      for (int dz = 0; dz < 2; dz++)
#pragma unroll
        for (int dy = 0; dy < 2; dy++) {
          const uint2 p = *reinterpret_cast<const uint2*>(&s0[2 * z + dz][2 * y + dy][2 * x]);
          w[child_id(0, dy, dz)] = p.x;
          w[child_id(1, dy, dz)] = p.y;
          any |= p.x | p.y;
        }
      uint32_t o[6];
      if (__any_sync(0xffffffffu, any != 0u)) {
        const ChildLoader ld{&s0[2 * z][2 * y][2 * x], TX, TX * TY};
        filter6_shared(w, any != 0u, ld, o);
      } else {
#pragma unroll
        for (int d = 0; d < 6; d++) o[d] = 0u;
      }
#pragma unroll
      for (int d = 0; d < 6; d++) sm.s1[d][z][y][x] = o[d];
    }
  }
  __syncthreads();

  // ---- level 2: 64 texels x 6 directions = 384 items, the direction is warp-uniform ----
#pragma unroll 1
  for (int u = t; u < 384; u += kFusedThreads) {
    const int d = u >> 6, tex = u & 63, x = tex & 7, y = (tex >> 3) & 3, z = tex >> 5;
    uint32_t w[8];
#pragma unroll
    for (int dz = 0; dz < 2; dz++)
#pragma unroll
      for (int dy = 0; dy < 2; dy++) {
        const uint2 p = *reinterpret_cast<const uint2*>(&sm.s1[d][2 * z + dz][2 * y + dy][2 * x]);
        w[child_id(0, dy, dz)] = p.x;
        w[child_id(1, dy, dz)] = p.y;
      }
    const ChildLoader ld{&sm.s1[d][2 * z][2 * y][2 * x], TX / 2, (TX / 2) * (TY / 2)};
    sm.s2[d][z][y][x] = filter1(w, d, ld);
  }
  store_tile_levels(&sm, a, ss, x0, y0, z0, t, true, false, false);   // overlaps the level-2 arithmetic of other warps
  __syncthreads();

  // ---- level 3: 8 texels x 6 directions ----
  if (t < 48) {
    const int d = t >> 3, tex = t & 7, x = tex & 3, y = tex >> 2;
    uint32_t w[8];
#pragma unroll
    for (int dz = 0; dz < 2; dz++)
#pragma unroll
      for (int dy = 0; dy < 2; dy++)
#pragma unroll
        for (int dx = 0; dx < 2; dx++) w[child_id(dx, dy, dz)] = sm.s2[d][dz][2 * y + dy][2 * x + dx];
    const ChildLoader ld{&sm.s2[d][0][2 * y][2 * x], TX / 4, (TX / 4) * (TY / 4)};
    sm.s3[d][y][x] = filter1(w, d, ld);
  }
  store_tile_levels(&sm, a, ss, x0, y0, z0, t, false, true, false);
  __syncthreads();
  store_tile_levels(&sm, a, ss, x0, y0, z0, t, false, false, true);
}

// The tiles of a CTA are blockIdx.x, blockIdx.x + gridDim.x, ...  In rounds of kListMax candidates the CTA first compacts the
// tiles it actually has to READ -- all of them when nothing is known about level 0; with the voxelizer's tile flags only the
// touched tiles and those whose outputs of the previous build are not zero yet (< 5 % of the tiles of the Cornell scene) --
// and then streams that list through the TMA ring.
__global__ void __launch_bounds__(kFusedThreads, 3)
mip_fused_kernel(const __grid_constant__ CUtensorMap tmap, const FusedArgs a) {
  extern __shared__ __align__(128) unsigned char fused_smem_raw[];
  FusedSmem& sm = *reinterpret_cast<FusedSmem*>(fused_smem_raw);
  const int t = threadIdx.x;
  if (t == 0) {
#pragma unroll
    for (int s = 0; s < kStages; s++) mbar_init(&sm.full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // make the initialised barriers visible to the async (TMA) proxy
  }
  StoreSlots ss;
  ss.l1_xq = t & 3; ss.l1_y = (t >> 2) & 7;
#pragma unroll
  for (int i = 0; i < 3; i++) { const int u = t + kFusedThreads * i; ss.l1_z[i] = (u >> 5) & 3; ss.l1_d[i] = u >> 7; }
  { const int u = (t - 64) & 127; ss.l2_xq = u & 1; ss.l2_y = (u >> 1) & 3; ss.l2_z = (u >> 3) & 1; ss.l2_d = min(u >> 4, 5); }

  for (uint32_t i = blockIdx.x * kFusedThreads + t; i < a.zero0_n; i += gridDim.x * kFusedThreads) a.zero0[i] = 0u;
  for (uint32_t i = blockIdx.x * kFusedThreads + t; i < a.zero1_n; i += gridDim.x * kFusedThreads) a.zero1[i] = 0u;
  const int n_my = ((int)blockIdx.x < a.n_tiles) ? (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  uint32_t it = 0;   // uses of the ring so far: stage = it % kStages, barrier parity = (it / kStages) & 1
  for (int cand0 = 0; cand0 < n_my; cand0 += kListMax) {
    __syncthreads();   // the previous round is done with the list (and the barriers are initialised)
    if (t == 0) sm.list_n = 0;
    __syncthreads();
    const int cand1 = min(cand0 + kListMax, n_my);
    for (int k = cand0 + t; k < cand1; k += kFusedThreads) {
      const uint32_t tile = blockIdx.x + (uint32_t)k * gridDim.x;
      const uint32_t zero = a.dense ? 0u : a.tile_zero[tile];
      const bool active = a.touched ? (a.touched[tile] != 0 || zero == 0u) : true;
      if (active) sm.list[atomicAdd(&sm.list_n, 1u)] = tile | (zero ? 0x80000000u : 0u);
    }
    __syncthreads();
    const int n = (int)sm.list_n;
    if (t == 0) {   // prologue: fill the ring
#pragma unroll
      for (int s = 0; s < kStages; s++) {
        if (s < n) {
          int bx, by, bz;
          tile_coords(sm.list[s] & 0x7FFFFFFFu, a, bx, by, bz);
          const int stage = (int)((it + (uint32_t)s) % kStages);
          tma_load_tile(&tmap, &sm.s0[stage][0][0][0], &sm.full[stage], bx * TX, by * TY, bz * TZ);
        }
      }
    }
    for (int i = 0; i < n; i++, it++) {
      const int stage = (int)(it % kStages);
      mbar_wait(&sm.full[stage], (it / kStages) & 1u);
      const uint32_t entry = sm.list[i];
      const uint32_t tile = entry & 0x7FFFFFFFu;
      int bx, by, bz;
      tile_coords(tile, a, bx, by, bz);
      process_tile(sm, sm.s0[stage], a, ss, tile, entry >> 31, bx, by, bz);
      __syncthreads();   // every read of this stage (and of s1/s2/s3) is done: the buffer can be refilled
      if (t == 0 && i + kStages < n) {
        int nx, ny, nz;
        tile_coords(sm.list[i + kStages] & 0x7FFFFFFFu, a, nx, ny, nz);
        tma_load_tile(&tmap, &sm.s0[stage][0][0][0], &sm.full[stage], nx * TX, ny * TY, nz * TZ);
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

// ---------------------------------------------------------------------------------------------
// Fused path, everything that follows the fused kernel in ONE wave of independent work (mip_tail_kernel):
//   blocks [0, n_sb)                 fold one super-block (64^3 voxels, or the whole grid when R = 32): its level-3 texels (linear
//                                    records of the fused kernel) -> levels 4, 5, 6 of the six directions + their occupancy bytes.
//                                    (bits of the occupied texels ORed into words the fused kernel zeroed).  The CTA that finishes last
//                                    builds levels 7..
//   blocks [n_sb, n_sb + word_blocks) dilation of levels 0-2 from their occupancy words, one thread per output word
//   the rest                          level 3: one WARP per output word, straight from the occupancy bytes -- lane j = texel 32 k + j:
//                                    plain bits (ballot of the byte) and dilated bits (ballot of the OR of the 2x2x2 bytes)
struct TailArgs {
  OccArgs occ;
  const uint32_t* rec3;        // level 3, records of six words
  uint32_t* rec_top;           // levels 5..: records at top_off[l] (scratch of this kernel)
  uint32_t top_off[VCT_MAX_LEVELS];
  uint8_t* occb;               // occupancy bytes of levels 3.. (level 3 written by the fused kernel, 4.. here)
  uint32_t* ticket;            // super-blocks finished
  SurfSet surf;
  int n_sb, sb3;               // super-blocks; level-3 texels per super-block side (8, or 4 when R = 32)
  int row_start[4];            // first block (relative to n_sb) of the dilated words of level 0, 1, 2 (and the end)
  int chunks[3];               // blocks per z-slice of those levels: a block = 256 consecutive words of ONE slice (no per-thread division by D)
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
    const ChildLoader ld{&src[d][2 * z][2 * y][2 * x], NS, NS * NS};
    const uint32_t o = filter1(w, d, ld);
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
      const uint32_t o = filter1(w, d);
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

__global__ void __launch_bounds__(256)
mip_tail_kernel(const TailArgs a) {
  __shared__ TailSmem sm;
  const int b = (int)blockIdx.x, t = (int)threadIdx.x;
  if (b < a.n_sb) {
    const int R = a.occ.R, N3 = R >> 3, sb3 = a.sb3, nsb = N3 / sb3;
    const int sx = b % nsb, sy = (b / nsb) % nsb, sz = b / (nsb * nsb);
    const uint8_t* ob3 = a.occb + a.occ.occb_off[3];
    for (int u = t; u < sb3 * sb3 * sb3 * 6; u += 256) {
      const int d = u % 6, tex = u / 6, x = tex % sb3, y = (tex / sb3) % sb3, z = tex / (sb3 * sb3);
      const size_t gi = ((size_t)(sz * sb3 + z) * N3 + (sy * sb3 + y)) * N3 + (sx * sb3 + x);
      sm.f[d][z][y][x] = a.rec3[gi * 6 + d];
      if (d == 0) sm.o3[z][y][x] = ob3[gi];
    }
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
    // ---- the last super-block to finish builds the levels above and the occupancy words of levels 4.. ----
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
    const int w = (lb - zz * a.chunks[level]) * 256 + t;                        // word of slice zz
    if (w < D * wpr) {
      const int yy = (int)(((float)w + 0.5f) * (1.0f / (float)wpr));           // exact: w < 2^16, wpr <= 33
      occ_dilate_word(a.occ, level, w - yy * wpr, yy, zz);
    }
    return;
  }
  const int warp = (lb0 - a.row_start[3]) * 8 + (t >> 5);
  if (warp < a.word_warps3) occ_words_from_bytes(a.occ, a.occb + a.occ.occb_off[3], 3, warp, t & 31);
}

// TMA descriptor of one level-0 buffer: 3-D u32 tensor R x R x R (x fastest), box = one 32 x 16 x 8 tile
static int make_base_tensor_map(uint32_t* base, int R, CUtensorMap* out) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    VCT_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
    if (!fn || qr != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return VCT_ERR_CUDA; }
    encode = (encode_fn)fn;
  }
  const cuuint64_t dims[3] = {(cuuint64_t)R, (cuuint64_t)R, (cuuint64_t)R};
  const cuuint64_t strides[2] = {(cuuint64_t)R * 4, (cuuint64_t)R * R * 4};   // bytes, dims 1 and 2
  const cuuint32_t box[3] = {TX, TY, TZ};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return VCT_ERR_CUDA; }
  return VCT_OK;
}

bool mip_fused_applies(int R, int levels) { return R >= 32 && levels >= 6; }

int launch_mipmap(vct_device* dev, vct_grid* g) {
  cudaStream_t s = dev->stream;
  const int R = g->R;
  OccArgs oa;
  memset(&oa, 0, sizeof oa);
  oa.R = R; oa.levels = g->levels; oa.src0 = g->base; oa.occb = g->occb;
  for (int l = 0; l < VCT_MAX_LEVELS; l++) { oa.occ[l] = g->occ[l]; oa.docc[l] = g->docc[l]; oa.occb_off[l] = g->occb_off[l]; }

  if (mip_fused_applies(R, g->levels)) {
    const int n_tiles = (R / TX) * (R / TY) * (R / TZ);
    const int buf = g->base == g->base_buf[1] ? 1 : 0;
    if (g->tmap_base_ptr[buf] != g->base) {   // descriptor of this level-0 buffer (built once)
      int rc = make_base_tensor_map(g->base, R, reinterpret_cast<CUtensorMap*>(g->tmap_storage[buf]));
      if (rc) return rc;
      g->tmap_base_ptr[buf] = g->base;
    }
    FusedArgs fa;
    memset(&fa, 0, sizeof fa);
    fa.R = R; fa.levels = g->levels; fa.n_tiles = n_tiles; fa.surf = g->surf;
    while ((TX << fa.log_tx) < R) fa.log_tx++;
    while ((TY << fa.log_ty) < R) fa.log_ty++;
    fa.occ0 = g->occ[0]; fa.occ1 = (uint16_t*)g->occ[1]; fa.occ2 = (uint8_t*)g->occ[2];
    fa.rec3 = g->rec3; fa.occb3 = g->occb + g->occb_off[3];
    fa.zero0 = g->occ_hi; fa.zero0_n = g->occ_hi_words; fa.zero1 = g->docc_all + g->docc_hi_off; fa.zero1_n = g->docc_hi_words;
    fa.tile_zero = g->tile_zero;
    fa.touched = (g->flags_valid && !g->external) ? g->tile_touched : nullptr;
    if (dev->debug_mip_dense) {   // measurement switch (vct_debug_set): the dense build, every tile read and written
      fa.touched = nullptr;
      fa.dense = 1;
    }
    if (!dev->mip_attr_set) {
      VCT_CUDA(cudaFuncSetAttribute(mip_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FusedSmem)));
      dev->mip_attr_set = true;
    }
    const int ctas = min(n_tiles, dev->prop.multiProcessorCount * 3);   // persistent: 3 CTAs of ~64 KB shared memory per SM
    mip_fused_kernel<<<ctas, kFusedThreads, sizeof(FusedSmem), s>>>(*reinterpret_cast<const CUtensorMap*>(g->tmap_storage[buf]), fa);
    TailArgs ta;
    memset(&ta, 0, sizeof ta);
    ta.occ = oa; ta.rec3 = g->rec3; ta.rec_top = g->rec_top; ta.occb = g->occb; ta.ticket = g->mip_counters; ta.surf = g->surf;
    for (int l = 0; l < VCT_MAX_LEVELS; l++) ta.top_off[l] = g->top_off[l];
    ta.sb3 = R >= 64 ? 8 : 4;
    ta.n_sb = R >= 64 ? (R / 64) * (R / 64) * (R / 64) : 1;
    int blocks = 0;
    for (int l = 0; l < 3; l++) {
      const int D = (R >> l) + 1;
      ta.row_start[l] = blocks;
      ta.chunks[l] = (D * occ_wpr(R >> l) + 255) / 256;
      blocks += D * ta.chunks[l];
    }
    ta.row_start[3] = blocks;
    ta.word_warps3 = occ_word_warps(R >> 3);
    mip_tail_kernel<<<ta.n_sb + blocks + (ta.word_warps3 + 7) / 8, 256, 0, s>>>(ta);
    VCT_CUDA(cudaGetLastError());
    return VCT_OK;
  }

  // ---- generic path (small grids / short chains): one launch per level ----
  for (int l = 0; l + 1 < g->levels; l++) {
    const int Ns = max(R >> l, 1), Nd = R >> (l + 1);
    if (Nd < 1) break;
    const size_t n = (size_t)Nd * Nd * Nd * 6;
    mip_generic_kernel<<<grid_for(n), 256, 0, s>>>(l == 0 ? g->base : nullptr, l == 0 ? 0 : g->surf.s[l], l == 0 ? 0 : g->surf.pitch[l], g->surf.s[l + 1],
                                                  g->surf.pitch[l + 1], Ns, Nd);
  }
  occ_bits_kernel<<<grid_for((size_t)R * R * R, 256, 148 * 8), 256, 0, s>>>(oa);
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
