// mipmap.cu -- six-direction anisotropic mip chain of the voxel grid.
//
// Replaces Renderer::filter() (src/renderer.cpp:283-314) + shader/mipmap.comp.  The reference
// dispatches one compute pass per level, each re-reading six full textures (and launching 8x more
// threads than texels).  Here the chain is built in TWO launches for the reference's 7 levels:
//   mip_fused_low_kernel   persistent CTAs stream 32x8x8 tiles of level 0 through a 4-stage TMA ring
//                          (cp.async.bulk.tensor.3d + mbarrier) -> levels 1, 2, 3 of all six directions
//                          and the occupancy bits of levels 0-2 (level 0 is read ONCE, not 6x).  Per
//                          round a CTA first compacts the tiles it has to read: with the voxelizer's
//                          tile flags, untouched tiles whose outputs are already zero are skipped.
//   mip_tail_kernel        one launch for everything that depends on the low kernel only: 8x8x8 tiles of
//                          level 3 -> levels 4, 5, 6; occupancy bits of the small levels; 2x2x2 dilation
//                          of the occupancy bits of every level
// so the dense DRAM traffic is the algorithmic minimum 4*R^3 (read) + 24*R^3*(1/8+1/64+...) (write), and
// the traffic of a running frame loop is proportional to the occupied tiles.
// Levels >= 1 are stored twice: records of six RGBA8 words per texel (direction-minor; software sampler,
// next mip level, downloads) and the stacked mipmapped array the texture units read (surface writes).
// All-zero child groups are skipped (exact: the filter of zeros is zero); mip_generic_kernel /
// occ_*_kernel cover grids the fused kernels do not (R < 32, fewer than 4 levels, more than 7).
//
// Arithmetic = oracle rules R5/R6 (built with -fmad=false): c/255.0f correctly rounded (multiply
// by 1/255 plus one exact Newton step, verified for all 256 inputs), blend f + (1-f.a)*b per
// mipmap.comp:40-43, sum of the four pairs in order, /4, rint(clamp*255) -> bit-exact vs the oracle.
#include <cuda.h>

#include <cstdlib>

#include "vct_internal.cuh"

namespace vct {

// children numbering of mipmap.comp:10-20: bit 2 = (x == 0), bit 1 = (y == 0), bit 0 = (z == 0)
// => child i sits at offset (x,y,z) = (!(i>>2&1), !(i>>1&1), !(i&1)).
// pairs[d][p] = {front, back} (mipmap.comp:59-98)
__device__ constexpr int kPairs[6][4][2] = {
    {{0, 4}, {1, 5}, {2, 6}, {3, 7}},  // -x
    {{4, 0}, {5, 1}, {6, 2}, {7, 3}},  // +x
    {{0, 2}, {1, 3}, {5, 7}, {4, 6}},  // -y
    {{2, 0}, {3, 1}, {7, 5}, {6, 4}},  // +y
    {{0, 1}, {2, 3}, {4, 5}, {6, 7}},  // -z
    {{1, 0}, {3, 2}, {5, 4}, {7, 6}},  // +z
};

// exact unorm8 -> float: fl(b / 255)
__device__ __forceinline__ float unorm(uint32_t word, int byte) {
  // byte -> float without I2F: insert it into the mantissa of 2^23 and subtract 2^23
  float b = __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7650 + byte)) - 8388608.0f;
  const float k = 0.003921568859368563f;  // fl(1/255)
  float q = b * k;
  float r = fmaf(-q, 255.0f, b);  // exact residual
  return fmaf(r, k, q);           // correctly rounded quotient
}
__device__ __forceinline__ void unpack4(uint32_t w, float c[4]) {
  c[0] = unorm(w, 0); c[1] = unorm(w, 1); c[2] = unorm(w, 2); c[3] = unorm(w, 3);
}
// rintf(clamp(v,0,1) * 255) via the 1.5*2^23 trick (ties-to-even, same as rintf)
__device__ __forceinline__ uint32_t to_unorm(float v) {
  float t = fminf(fmaxf(v, 0.0f), 1.0f) * 255.0f;
  return __float_as_uint(t + 12582912.0f) & 0xFFu;
}

// one direction of one parent texel from its 8 unpacked children
template <int D>
__device__ __forceinline__ uint32_t filter_dir(const float (&c)[8][4]) {
  uint32_t out = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < 4; p++) {
      const int f = kPairs[D][p][0], b = kPairs[D][p][1];
      float v = c[f][k] + ((1.0f - c[f][3]) * c[b][k]);
      s = p == 0 ? v : s + v;
    }
    out |= to_unorm(s * 0.25f) << (8 * k);
  }
  return out;
}

__device__ __forceinline__ uint32_t filter_dir_dyn(const float (&c)[8][4], int d) {
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

// child index i of mipmap.comp for local offsets (dx,dy,dz)
__device__ __forceinline__ constexpr int child_id(int dx, int dy, int dz) { return ((dx ^ 1) << 2) | ((dy ^ 1) << 1) | (dz ^ 1); }

// ---------------------------------------------------------------------------------------------
// generic fallback: one thread per (destination texel, direction)
__global__ void mip_generic_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int Ns, int Nd, int src_is_base, SurfSet surf, int dst_level) {
  const size_t n = (size_t)Nd * Nd * Nd * 6;
  for (size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x; u < n; u += (size_t)gridDim.x * blockDim.x) {
    const int d = (int)(u % 6);
    size_t tex = u / 6;
    const int x = (int)(tex % Nd), y = (int)((tex / Nd) % Nd), z = (int)(tex / ((size_t)Nd * Nd));
    float c[8][4];
    uint32_t any = 0;
#pragma unroll
    for (int dz = 0; dz < 2; dz++)
#pragma unroll
      for (int dy = 0; dy < 2; dy++)
#pragma unroll
        for (int dx = 0; dx < 2; dx++) {
          size_t si = ((size_t)(2 * z + dz) * Ns + (2 * y + dy)) * Ns + (2 * x + dx);
          uint32_t w = src_is_base ? src[si] : src[si * 6 + d];
          any |= w;
          unpack4(w, c[child_id(dx, dy, dz)]);
        }
    const uint32_t o = any ? filter_dir_dyn(c, d) : 0u;
    dst[u] = o;
    surf_write(surf, d, dst_level, o, x * 4, y, z);
  }
}

// ---------------------------------------------------------------------------------------------
// fused levels 0 -> 1,2,3.  Tile = 32 x 8 x 8 level-0 texels (8 KB).  PERSISTENT kernel: each CTA walks tiles
// blockIdx.x, blockIdx.x + gridDim.x, ...; the tiles are fetched by TMA (cp.async.bulk.tensor.3d, one instruction per
// tile, issued by one thread) into a ring of kLowStages shared-memory buffers and signalled through mbarriers, so that
// kLowStages-1 tiles per CTA are in flight while one is being reduced -- the HBM stream never waits for the arithmetic.
constexpr int TX = 32, TY = 8, TZ = 8;
constexpr int kLowStages = 4;
constexpr int kLowListMax = 1024;              // tiles of one CTA examined per round
constexpr uint32_t kLowTileBytes = TX * TY * TZ * 4;

struct LowArgs {
  uint32_t *l1, *l2, *l3;
  uint32_t* occ0; uint16_t* occ1; uint8_t* occ2;
  int R, n_tiles, log_tiles_x, log_tiles_y;
  uint8_t* tile_zero;   // per tile: 1 = every output of this tile (levels 1-3, both copies, occupancy bits) is known to be zero
  const uint8_t* touched;   // per tile: the voxelizer wrote into it since the last clear (nullptr: unknown, every tile is read)
  SurfSet surf;
};

struct LowSmem {
  uint32_t s0[kLowStages][TZ][TY][TX];         // 4 x 8 KB, TMA destinations (128-byte aligned)
  uint32_t s1[TZ / 2][TY / 2][TX / 2][6];      // 6 KB
  uint32_t s2[TZ / 4][TY / 4][TX / 4][6];      // 768 B
  uint32_t s3[TX / 8][6];                      // 96 B
  uint32_t occ_l1[8];                          // per warp of the level-1 step: 8 bits = (x pairs) of its two rows with a non-zero voxel below
  uint32_t list[kLowListMax];                  // the tiles this CTA has to read (bit 31: tile_zero[] of the tile)
  uint32_t list_n;
  unsigned long long full[kLowStages];         // mbarriers: "tile landed"
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
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(kLowTileBytes) : "memory");
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
               "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
               : "memory");
}

// R is a power of two (vct_grid_create), so are the tile counts: shifts instead of integer divisions
__device__ __forceinline__ void low_tile_coords(int tile, int log_tiles_x, int log_tiles_y, int& bx, int& by, int& bz) {
  bx = tile & ((1 << log_tiles_x) - 1);
  by = (tile >> log_tiles_x) & ((1 << log_tiles_y) - 1);
  bz = tile >> (log_tiles_x + log_tiles_y);
}

// Writes one level's share of a tile from shared memory (or zeros when src == nullptr) with 16-byte stores:
// ROWS rows of NX texels; records: a row is NX*24 contiguous bytes; arrays: per direction NX*4 contiguous bytes.
template <int NX, int ROWS_Y, int ROWS_Z>
__device__ __forceinline__ void low_store_level(const uint32_t* __restrict__ src /* [ROWS_Z][ROWS_Y][NX][6] */, uint32_t* __restrict__ rec, const SurfSet& surf, int level,
                                                int N, int x1, int y1, int z1, int t) {
  constexpr int kRowQuads = NX * 6 / 4;            // uint4 per record row
  constexpr int kRecQuads = ROWS_Y * ROWS_Z * kRowQuads;
  for (int u = t; u < kRecQuads; u += 256) {
    const int row = u / kRowQuads, q = u % kRowQuads, y = row % ROWS_Y, z = row / ROWS_Y;
    const uint4 v = src ? *reinterpret_cast<const uint4*>(src + (size_t)row * NX * 6 + 4 * q) : make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(rec + (((size_t)(z1 + z) * N + (y1 + y)) * N + x1) * 6 + 4 * q) = v;
  }
  constexpr int kXQ = NX / 4;                       // 4 texels (16 bytes) per surface store
  constexpr int kSurfQuads = ROWS_Y * ROWS_Z * 6 * kXQ;
  for (int u = t; u < kSurfQuads; u += 256) {
    const int xq = u % kXQ, d = (u / kXQ) % 6, row = u / (kXQ * 6), y = row % ROWS_Y, z = row / ROWS_Y;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (src) {
      const uint32_t* p = src + ((size_t)row * NX + 4 * xq) * 6 + d;
      v = make_uint4(p[0], p[6], p[12], p[18]);
    }
    surf_write(surf, d, level, v, (x1 + 4 * xq) * 4, y1 + y, z1 + z);
  }
}

// reduces one staged tile; every thread of the CTA calls it (contains barriers)
__device__ __forceinline__ void low_process_tile(const uint32_t (&s0)[TZ][TY][TX], uint32_t (&s1)[TZ / 2][TY / 2][TX / 2][6],
                                                 uint32_t (&s2)[TZ / 4][TY / 4][TX / 4][6], uint32_t (&s3)[TX / 8][6], uint32_t (&occ_l1)[8], const LowArgs& a, int tile,
                                                 uint32_t known_zero, int bx, int by, int bz) {
  const int R = a.R;
  uint32_t* __restrict__ occ0 = a.occ0; uint16_t* __restrict__ occ1 = a.occ1; uint8_t* __restrict__ occ2 = a.occ2;
  const int x0 = bx * TX, y0 = by * TY, z0 = bz * TZ;
  const int t = threadIdx.x;
  const int N1 = R >> 1, N2 = R >> 2, N3 = R >> 3;

  // ---- "is the tile empty": 64 rows of 128 B, two 16-byte shared loads per thread ----
  uint4 v0[2];
#pragma unroll
  for (int k = 0; k < 2; k++) v0[k] = *reinterpret_cast<const uint4*>(&s0[((t >> 3) + 32 * k) >> 3][((t >> 3) + 32 * k) & 7][4 * (t & 7)]);
  const uint32_t any0 = (v0[0].x | v0[0].y | v0[0].z | v0[0].w) | (v0[1].x | v0[1].y | v0[1].z | v0[1].w);
  const int tile_nonzero = __syncthreads_or((int)(any0 != 0u));
  // empty now and every output (incl. the occupancy words) known to be zero from the previous build: nothing to write.
  // The usual case: < 1 % of the grid is occupied and the occupied set moves little between frames.
  if (!tile_nonzero && known_zero) return;
  // ---- occupancy bits of level 0: 8 consecutive lanes hold the 32 voxels of a row ----
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const int row = (t >> 3) + 32 * k, quad = t & 7;
    const int y = row & 7, z = row >> 3;
    const uint4 v = v0[k];
    uint32_t bits = ((v.x != 0u) | (v.y != 0u) << 1 | (v.z != 0u) << 2 | (v.w != 0u) << 3) << (4 * quad);
    bits |= __shfl_xor_sync(0xffffffffu, bits, 1);
    bits |= __shfl_xor_sync(0xffffffffu, bits, 2);
    bits |= __shfl_xor_sync(0xffffffffu, bits, 4);
    if (quad == 0) occ0[((size_t)(z0 + z) * R + (y0 + y)) * (R / 32) + bx] = bits;
  }

  if (!tile_nonzero) {
    // empty tile: every output of this tile is zero (exact: the filter of zeros is zero)
    if (t == 0) a.tile_zero[tile] = 1;
    if (t < 16) occ1[((size_t)(z0 / 2 + (t >> 2)) * N1 + (y0 / 2 + (t & 3))) * (N1 / 16) + bx] = 0;
    if (t >= 32 && t < 36) occ2[((size_t)(z0 / 4 + ((t >> 1) & 1)) * N2 + (y0 / 4 + (t & 1))) * (N2 / 8) + bx] = 0;
    low_store_level<TX / 2, TY / 2, TZ / 2>(nullptr, a.l1, a.surf, 1, N1, x0 / 2, y0 / 2, z0 / 2, t);
    low_store_level<TX / 4, TY / 4, TZ / 4>(nullptr, a.l2, a.surf, 2, N2, x0 / 4, y0 / 4, z0 / 4, t);
    low_store_level<TX / 8, 1, 1>(nullptr, a.l3, a.surf, 3, N3, x0 / 8, y0 / 8, z0 / 8, t);
    return;
  }

  if (t == 0) a.tile_zero[tile] = 0;
  // ---- level 1: one texel per thread, six directions ----
  {
    const int x = t & 15, y = (t >> 4) & 3, z = t >> 6;
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
    uint32_t o[6] = {0u, 0u, 0u, 0u, 0u, 0u};
    if (any) {
      float c[8][4];
#pragma unroll
      for (int i = 0; i < 8; i++) unpack4(w[i], c[i]);
      o[0] = filter_dir<0>(c); o[1] = filter_dir<1>(c); o[2] = filter_dir<2>(c);
      o[3] = filter_dir<3>(c); o[4] = filter_dir<4>(c); o[5] = filter_dir<5>(c);
    }
    uint2* sp = reinterpret_cast<uint2*>(&s1[z][y][x][0]);
    sp[0] = make_uint2(o[0], o[1]); sp[1] = make_uint2(o[2], o[3]); sp[2] = make_uint2(o[4], o[5]);
    // occupancy of a texel of level >= 1 = "a voxel of its level-0 support is non-zero" (a superset of "the texel is non-zero": a
    // filtered value can round to zero), so that the bits of a level are the OR of the 8 child bits -- the tracer relies on that.
    // A warp holds two rows of 16 texels.
    const uint32_t bal = __ballot_sync(0xffffffffu, any != 0u);
    if (x == 0) occ1[((size_t)(z0 / 2 + z) * N1 + (y0 / 2 + y)) * (N1 / 16) + bx] = (uint16_t)(bal >> (16 * (y & 1)));
    if ((t & 31) == 0) occ_l1[t >> 5] = occ_pair_or((bal | (bal >> 16)) & 0xFFFFu);   // warp w: z = w >> 1, level-2 row y = w & 1
  }
  __syncthreads();

  // ---- level 2: 32 texels x 6 directions = 192 threads ----
  if (t < 192) {
    const int d = t % 6, tex = t / 6, x = tex & 7, y = (tex >> 3) & 1, z = tex >> 4;
    float c[8][4];
    uint32_t any = 0;
#pragma unroll
    for (int dz = 0; dz < 2; dz++)
#pragma unroll
      for (int dy = 0; dy < 2; dy++)
#pragma unroll
        for (int dx = 0; dx < 2; dx++) {
          uint32_t w = s1[2 * z + dz][2 * y + dy][2 * x + dx][d];
          any |= w;
          unpack4(w, c[child_id(dx, dy, dz)]);
        }
    s2[z][y][x][d] = any ? filter_dir_dyn(c, d) : 0u;
  }
  low_store_level<TX / 2, TY / 2, TZ / 2>(&s1[0][0][0][0], a.l1, a.surf, 1, N1, x0 / 2, y0 / 2, z0 / 2, t);   // overlaps the level-2 arithmetic of other warps
  __syncthreads();

  if (t >= 192 && t < 196) {  // level-2 occupancy: one byte per row of 8 texels
    const int y = t & 1, z = (t >> 1) & 1;
    const uint32_t bits = occ_l1[4 * z + y] | occ_l1[4 * z + 2 + y];
    occ2[((size_t)(z0 / 4 + z) * N2 + (y0 / 4 + y)) * (N2 / 8) + bx] = (uint8_t)bits;
  }
  // ---- level 3: 4 texels x 6 directions ----
  if (t < 24) {
    const int d = t % 6, x = t / 6;
    float c[8][4];
    uint32_t any = 0;
#pragma unroll
    for (int dz = 0; dz < 2; dz++)
#pragma unroll
      for (int dy = 0; dy < 2; dy++)
#pragma unroll
        for (int dx = 0; dx < 2; dx++) {
          uint32_t w = s2[dz][dy][2 * x + dx][d];
          any |= w;
          unpack4(w, c[child_id(dx, dy, dz)]);
        }
    s3[x][d] = any ? filter_dir_dyn(c, d) : 0u;
  }
  low_store_level<TX / 4, TY / 4, TZ / 4>(&s2[0][0][0][0], a.l2, a.surf, 2, N2, x0 / 4, y0 / 4, z0 / 4, t);
  __syncthreads();
  low_store_level<TX / 8, 1, 1>(&s3[0][0], a.l3, a.surf, 3, N3, x0 / 8, y0 / 8, z0 / 8, t);
}

// The tiles of a CTA are blockIdx.x, blockIdx.x + gridDim.x, ...  In rounds of kLowListMax candidates the CTA first compacts the
// tiles it actually has to READ -- all of them when nothing is known about level 0; with the voxelizer's tile flags only the
// touched tiles and those whose outputs of the previous build are not zero yet (< 5 % of the tiles of the Cornell scene: the
// level is neither read nor written elsewhere) -- and then streams that list through the TMA ring.
__global__ void __launch_bounds__(256, 4)
mip_fused_low_kernel(const __grid_constant__ CUtensorMap tmap, const LowArgs a) {
  extern __shared__ __align__(128) unsigned char low_smem_raw[];
  LowSmem& sm = *reinterpret_cast<LowSmem*>(low_smem_raw);
  const int t = threadIdx.x;
  if (t == 0) {
#pragma unroll
    for (int s = 0; s < kLowStages; s++) mbar_init(&sm.full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // make the initialised barriers visible to the async (TMA) proxy
  }
  const int n_my = ((int)blockIdx.x < a.n_tiles) ? (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  uint32_t it = 0;   // uses of the ring so far: stage = it % kLowStages, barrier parity = (it / kLowStages) & 1
  for (int cand0 = 0; cand0 < n_my; cand0 += kLowListMax) {
    __syncthreads();   // the previous round is done with the list (and the barriers are initialised)
    if (t == 0) sm.list_n = 0;
    __syncthreads();
    const int cand1 = min(cand0 + kLowListMax, n_my);
    for (int k = cand0 + t; k < cand1; k += 256) {
      const uint32_t tile = blockIdx.x + (uint32_t)k * gridDim.x;
      const uint32_t zero = a.tile_zero[tile];
      const bool active = a.touched ? (a.touched[tile] != 0 || zero == 0u) : true;
      if (active) sm.list[atomicAdd(&sm.list_n, 1u)] = tile | (zero ? 0x80000000u : 0u);
    }
    __syncthreads();
    const int n = (int)sm.list_n;
    if (t == 0) {   // prologue: fill the ring
#pragma unroll
      for (int s = 0; s < kLowStages; s++) {
        if (s < n) {
          int bx, by, bz;
          low_tile_coords((int)(sm.list[s] & 0x7FFFFFFFu), a.log_tiles_x, a.log_tiles_y, bx, by, bz);
          const int stage = (int)((it + (uint32_t)s) % kLowStages);
          tma_load_tile(&tmap, &sm.s0[stage][0][0][0], &sm.full[stage], bx * TX, by * TY, bz * TZ);
        }
      }
    }
    for (int i = 0; i < n; i++, it++) {
      const int stage = (int)(it % kLowStages);
      mbar_wait(&sm.full[stage], (it / kLowStages) & 1u);
      const uint32_t entry = sm.list[i];
      const int tile = (int)(entry & 0x7FFFFFFFu);
      int bx, by, bz;
      low_tile_coords(tile, a.log_tiles_x, a.log_tiles_y, bx, by, bz);
      low_process_tile(sm.s0[stage], sm.s1, sm.s2, sm.s3, sm.occ_l1, a, tile, entry >> 31, bx, by, bz);
      __syncthreads();   // every read of this stage (and of s1/s2) is done: the buffer can be refilled
      if (t == 0 && i + kLowStages < n) {
        low_tile_coords((int)(sm.list[i + kLowStages] & 0x7FFFFFFFu), a.log_tiles_x, a.log_tiles_y, bx, by, bz);
        tma_load_tile(&tmap, &sm.s0[stage][0][0][0], &sm.full[stage], bx * TX, by * TY, bz * TZ);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// fused levels 3 -> 4,5,6.  Tile = 8^3 level-3 records, 256 threads.
struct HighSmem {
  uint32_t s3[8][8][8][6];  // 12 KB
  uint32_t s4[4][4][4][6];
  uint32_t s5[2][2][2][6];
};
__device__ __forceinline__ void mip_high_tile(HighSmem& sm, int block, const uint32_t* __restrict__ l3, uint32_t* __restrict__ l4, uint32_t* __restrict__ l5,
                                              uint32_t* __restrict__ l6, int N3, const SurfSet& surf) {
  uint32_t (&s3)[8][8][8][6] = sm.s3;
  uint32_t (&s4)[4][4][4][6] = sm.s4;
  uint32_t (&s5)[2][2][2][6] = sm.s5;
  const int tiles = N3 / 8;
  const int bx = block % tiles, by = (block / tiles) % tiles, bz = block / (tiles * tiles);
  const int t = threadIdx.x;
  // 64 rows (z,y) of 48 words
  for (int u = t; u < 64 * 48; u += 256) {
    const int row = u / 48, wdx = u % 48, y = row & 7, z = row >> 3;
    (&s3[z][y][0][0])[wdx] = l3[(((size_t)(bz * 8 + z) * N3 + (by * 8 + y)) * N3 + bx * 8) * 6 + wdx];
  }
  __syncthreads();
  const int N4 = N3 / 2, N5 = N3 / 4, N6 = N3 / 8;
  for (int u = t; u < 64 * 6; u += 256) {
    const int d = u % 6, tex = u / 6, x = tex & 3, y = (tex >> 2) & 3, z = tex >> 4;
    float c[8][4];
    uint32_t any = 0;
#pragma unroll
    for (int dz = 0; dz < 2; dz++)
#pragma unroll
      for (int dy = 0; dy < 2; dy++)
#pragma unroll
        for (int dx = 0; dx < 2; dx++) {
          uint32_t w = s3[2 * z + dz][2 * y + dy][2 * x + dx][d];
          any |= w;
          unpack4(w, c[child_id(dx, dy, dz)]);
        }
    const uint32_t o = any ? filter_dir_dyn(c, d) : 0u;
    l4[(((size_t)(bz * 4 + z) * N4 + (by * 4 + y)) * N4 + (bx * 4 + x)) * 6 + d] = o;
    surf_write(surf, d, 4, o, (bx * 4 + x) * 4, by * 4 + y, bz * 4 + z);
    s4[z][y][x][d] = o;
  }
  __syncthreads();
  if (t < 48) {
    const int d = t % 6, tex = t / 6, x = tex & 1, y = (tex >> 1) & 1, z = tex >> 2;
    float c[8][4];
    uint32_t any = 0;
#pragma unroll
    for (int dz = 0; dz < 2; dz++)
#pragma unroll
      for (int dy = 0; dy < 2; dy++)
#pragma unroll
        for (int dx = 0; dx < 2; dx++) {
          uint32_t w = s4[2 * z + dz][2 * y + dy][2 * x + dx][d];
          any |= w;
          unpack4(w, c[child_id(dx, dy, dz)]);
        }
    const uint32_t o = any ? filter_dir_dyn(c, d) : 0u;
    l5[(((size_t)(bz * 2 + z) * N5 + (by * 2 + y)) * N5 + (bx * 2 + x)) * 6 + d] = o;
    surf_write(surf, d, 5, o, (bx * 2 + x) * 4, by * 2 + y, bz * 2 + z);
    s5[z][y][x][d] = o;
  }
  __syncthreads();
  if (t < 6) {
    const int d = t;
    float c[8][4];
    uint32_t any = 0;
#pragma unroll
    for (int dz = 0; dz < 2; dz++)
#pragma unroll
      for (int dy = 0; dy < 2; dy++)
#pragma unroll
        for (int dx = 0; dx < 2; dx++) {
          uint32_t w = s5[dz][dy][dx][d];
          any |= w;
          unpack4(w, c[child_id(dx, dy, dz)]);
        }
    const uint32_t o = any ? filter_dir_dyn(c, d) : 0u;
    l6[(((size_t)bz * N6 + by) * N6 + bx) * 6 + d] = o;
    surf_write(surf, d, 6, o, bx * 4, by, bz);
  }
}

__global__ void __launch_bounds__(256)
mip_fused_high_kernel(const uint32_t* __restrict__ l3, uint32_t* __restrict__ l4, uint32_t* __restrict__ l5, uint32_t* __restrict__ l6, int N3, const SurfSet surf) {
  __shared__ HighSmem sm;
  mip_high_tile(sm, (int)blockIdx.x, l3, l4, l5, l6, N3, surf);
}

// ---------------------------------------------------------------------------------------------
// occupancy bits
struct OccArgs {
  const uint32_t* src[VCT_MAX_LEVELS];  // level 0: base words, level >= 1: 6-word records
  uint32_t* occ[VCT_MAX_LEVELS];
  uint32_t* docc[VCT_MAX_LEVELS];
  int R, first_level;
};

__global__ void __launch_bounds__(256)
occ_bits_kernel(const OccArgs a) {   // level 0 from the voxel words (only when the fused kernel did not run)
  const size_t N = (size_t)a.R, n = N * N * N;
  const uint32_t* src = a.src[0];
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

// the small levels (N <= 32) depend on each other and are tiny: ONE CTA walks them in order
__device__ __forceinline__ void occ_reduce_small(const OccArgs& a, int first_level, int levels) {
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
    __syncthreads();   // the next level (and the dilation below) reads what this CTA just wrote
  }
}
__global__ void __launch_bounds__(1024)
occ_reduce_kernel(const OccArgs a, int first_level, int levels) { occ_reduce_small(a, first_level, levels); }

// 32 occupancy bits of row (y,z) starting at x = 32*k; rows outside the level read as zero
__device__ __forceinline__ uint32_t occ_row_bits(const uint32_t* __restrict__ occ, int N, int y, int z, int k) {
  if ((unsigned)y >= (unsigned)N || (unsigned)z >= (unsigned)N || k < 0 || k * 32 >= N) return 0u;
  const size_t flat = ((size_t)z * N + y) * N + (size_t)k * 32;
  if (N >= 32) return occ[flat >> 5];
  return (occ[flat >> 5] >> (flat & 31)) & ((1u << N) - 1u);
}

// dilation: docc bit (x+1,y+1,z+1) = OR of occ over [x,x+1]x[y,y+1]x[z,z+1].  One thread per output ROW (yy,zz) of a level:
// walks the row's words carrying the top bit of the previous word.  PLAIN loads: in the merged tail kernel the small levels are
// written by the same launch.
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

// grid: x = rows of one z-slice, y = z-slice, z = level
__global__ void __launch_bounds__(128)
occ_dilate_kernel(const OccArgs a) {
  const int level = (int)blockIdx.z;
  const int D = (a.R >> level) + 1;
  const int zz = (int)blockIdx.y, yy = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (zz >= D || yy >= D) return;
  occ_dilate_row(a, level, yy, zz);
}

// ---------------------------------------------------------------------------------------------
// Everything that follows the fused low kernel, in ONE launch (three independent jobs that each depend on the low kernel only;
// as three launches they cost 13 + 6 + 8 us at 256^3, most of it launch and tail latency):
//   blocks [0, n_high)            levels 3 -> 4, 5, 6 of the colour pyramid (mip_high_tile)
//   block  n_high                 occupancy bits of the small levels (N <= 32) + their dilation
//   blocks (n_high, ...)          dilation of the large levels, 256 rows per block
struct TailArgs {
  const uint32_t* l3; uint32_t *l4, *l5, *l6;
  int N3, n_high;
  OccArgs occ;
  int small_first, levels;      // the small levels are small_first .. levels-1
  int dil_start[VCT_MAX_LEVELS + 1];   // first block (relative to n_high + 1) of the dilation of each large level 0 .. small_first-1
  SurfSet surf;
};

__global__ void __launch_bounds__(256)
mip_tail_kernel(const TailArgs a) {
  __shared__ HighSmem sm;
  const int b = (int)blockIdx.x;
  if (b < a.n_high) {
    mip_high_tile(sm, b, a.l3, a.l4, a.l5, a.l6, a.N3, a.surf);
  } else if (b == a.n_high) {
    occ_reduce_small(a.occ, a.small_first, a.levels);
    for (int level = a.small_first; level < a.levels; level++) {
      const int D = (a.occ.R >> level) + 1;
      for (int r = threadIdx.x; r < D * D; r += blockDim.x) occ_dilate_row(a.occ, level, r % D, r / D);
    }
  } else {
    const int rb = b - a.n_high - 1;
    int level = 0;
    while (level + 1 < a.small_first && rb >= a.dil_start[level + 1]) level++;
    const int D = (a.occ.R >> level) + 1;
    const int r = (rb - a.dil_start[level]) * 256 + (int)threadIdx.x;
    if (r < D * D) occ_dilate_row(a.occ, level, r % D, r / D);
  }
}

// TMA descriptor of one level-0 buffer: 3-D u32 tensor R x R x R (x fastest), box = one 32 x 8 x 8 tile
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

int launch_mipmap(vct_device* dev, vct_grid* g) {
  cudaStream_t s = dev->stream;
  const int R = g->R;
  int level = 0;  // highest level already built
  int occ_from_data = 0;
  if (g->levels >= 4 && R % 32 == 0 && R >= 32) {
    const int n_tiles = (R / TX) * (R / TY) * (R / TZ);
    const int buf = g->base == g->base_buf[1] ? 1 : 0;
    if (g->tmap_base_ptr[buf] != g->base) {   // descriptor of this level-0 buffer (built once)
      int rc = make_base_tensor_map(g->base, R, reinterpret_cast<CUtensorMap*>(g->tmap_storage[buf]));
      if (rc) return rc;
      g->tmap_base_ptr[buf] = g->base;
    }
    LowArgs la;
    la.l1 = g->lvl[1]; la.l2 = g->lvl[2]; la.l3 = g->lvl[3];
    la.occ0 = g->occ[0]; la.occ1 = (uint16_t*)g->occ[1]; la.occ2 = (uint8_t*)g->occ[2];
    la.R = R; la.n_tiles = n_tiles; la.surf = g->surf;
    la.log_tiles_x = 0; la.log_tiles_y = 0;
    while ((TX << la.log_tiles_x) < R) la.log_tiles_x++;
    while ((TY << la.log_tiles_y) < R) la.log_tiles_y++;
    if (!g->tile_zero) {
      VCT_CUDA(cudaMalloc(&g->tile_zero, (size_t)n_tiles));
      VCT_CUDA(cudaMemsetAsync(g->tile_zero, 0, (size_t)n_tiles, s));   // unknown: the first build writes everything
    }
    la.tile_zero = g->tile_zero;
    la.touched = (g->flags_valid && !g->external) ? g->tile_touched : nullptr;
    if (const char* dense = getenv("VCT_MIP_DENSE"); dense && dense[0] == '1') {   // measurement switch: the dense build (every tile read and written)
      la.touched = nullptr;
      VCT_CUDA(cudaMemsetAsync(g->tile_zero, 0, (size_t)n_tiles, s));
    }
    const int ctas = min(n_tiles, dev->prop.multiProcessorCount * 4);   // persistent: 4 CTAs of 40 KB shared memory per SM
    mip_fused_low_kernel<<<ctas, 256, sizeof(LowSmem), s>>>(*reinterpret_cast<const CUtensorMap*>(g->tmap_storage[buf]), la);
    level = 3;
    occ_from_data = 3;  // levels 0..2 got their occupancy bits from the fused kernel
  }
  OccArgs oa;
  oa.R = R; oa.first_level = occ_from_data;
  for (int l = 0; l < VCT_MAX_LEVELS; l++) { oa.src[l] = l == 0 ? g->base : g->lvl[l]; oa.occ[l] = g->occ[l]; oa.docc[l] = g->docc[l]; }
  if (occ_from_data == 0) {
    occ_bits_kernel<<<grid_for((size_t)R * R * R, 256, 148 * 8), 256, 0, s>>>(oa);
    oa.first_level = 1;
  }
  // occupancy bits of the large levels (N >= 64) that the fused kernel did not produce: one grid-wide launch each
  int small_first = oa.first_level;
  for (; small_first < g->levels && (R >> small_first) >= 64; small_first++) {
    int logN = 0;
    while ((1 << logN) < (R >> small_first)) logN++;
    const size_t n_words = (size_t)1 << (3 * logN - 5);
    occ_reduce_level_kernel<<<grid_for(n_words), 256, 0, s>>>(g->occ[small_first - 1], g->occ[small_first], logN);
  }
  bool occ_done = false;
  if (level == 3 && g->levels >= 7 && (R >> 3) % 8 == 0) {
    // levels 3 -> 6, the small occupancy levels and every dilation in one launch
    const int tiles = (R >> 3) / 8;
    TailArgs ta;
    ta.l3 = g->lvl[3]; ta.l4 = g->lvl[4]; ta.l5 = g->lvl[5]; ta.l6 = g->lvl[6];
    ta.N3 = R >> 3; ta.n_high = tiles * tiles * tiles;
    ta.occ = oa; ta.small_first = small_first; ta.levels = g->levels; ta.surf = g->surf;
    int blocks = 0;
    for (int l = 0; l <= VCT_MAX_LEVELS; l++) ta.dil_start[l] = 0;
    for (int l = 0; l < small_first; l++) {
      const int D = (R >> l) + 1;
      ta.dil_start[l] = blocks;
      blocks += (D * D + 255) / 256;
    }
    ta.dil_start[small_first] = blocks;
    mip_tail_kernel<<<ta.n_high + 1 + blocks, 256, 0, s>>>(ta);
    level = 6;
    occ_done = true;
  }
  for (int l = level; l + 1 < g->levels; l++) {
    const int Ns = max(R >> l, 1), Nd = R >> (l + 1);
    if (Nd < 1) break;
    const size_t n = (size_t)Nd * Nd * Nd * 6;
    const int blocks = (int)grid_for(n);
    mip_generic_kernel<<<blocks, 256, 0, s>>>(l == 0 ? g->base : g->lvl[l], g->lvl[l + 1], Ns, Nd, l == 0, g->surf, l + 1);
  }
  if (!occ_done) {   // occupancy masks for the cone tracer's zero-footprint skip
    if (small_first < g->levels) occ_reduce_kernel<<<1, 1024, 0, s>>>(oa, small_first, g->levels);
    occ_dilate_kernel<<<dim3((R + 1 + 127) / 128, R + 1, g->levels), 128, 0, s>>>(oa);
  }
  VCT_CUDA(cudaGetLastError());
  return VCT_OK;
}

}  // namespace vct
