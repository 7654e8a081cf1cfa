// cone_trace.cu -- per-pixel diffuse / specular / shadow / refraction cone tracing over the G-buffer.
//
// Replaces shader/voxel_cone_tracing.frag (src/renderer.cpp:355-390 binds it).  Not a port of the
// fragment shader's one-thread-per-fragment loop.  Three launches:
//   tile_list_kernel   compacts the 8x4 screen tiles that contain at least one shaded pixel (on the G-buffer stream)
//   cone_kernel_fast   ONE WARP = one live tile x one job: all 32 lanes march the same cone of neighbouring pixels
//                      (same aperture, same step and LOD sequence -> coherent fetches, no divergence between cone
//                      types; four consecutive lanes = a 2x2 pixel block = one TEX quad).  Jobs: every cone slot on its
//                      own (9 diffuse, 1 specular, 1 refraction, 1 shadow per light), or -- frames with >= 32 k tiles
//                      per GPU -- all diffuse cones of the tile in one warp, which stores their sum.  Warps are
//                      independent (no barrier), the long jobs are scheduled first, results go to a [job][pixel] buffer.
//   shade_kernel       per pixel: Blinn-Phong / mix of main() (voxel_cone_tracing.frag:246-275).
//
// textureLod has two evaluators (vct_trace_params_t.sampler):
//   VCT_SAMPLER_TEX    levels >= 1 through the texture units from ONE mipmapped array that stacks the six directional
//                      volumes along z (GridView in vct_internal.cuh), level 0 in software (shared by the three
//                      directions).  The benchmarked path; frame within 2/255, PSNR > 64 dB of the oracle.
//   VCT_SAMPLER_FP32   software trilinear + mip-linear with fp32 weights from point-sampled texels of the array (rule R7 of the
//                      oracle exactly): trilinear in floor(lod) and floor(lod)+1, CLAMP_TO_BORDER with a zero border.
// Result-preserving savings over the literal shader (every skipped term is exactly zero in the reference): level 0 is
// fetched once for the three directions (all six level-0 textures are identical), a level whose filter footprint is empty
// (dilated occupancy bits) or wholly outside the grid is skipped, a direction of weight 0 is skipped, a fetch that needs one
// level goes through the nearest-mip texture object, and a cone stops once it has left the border-padded cube for good.
// cone_kernel<COUNT, TEX> is the literal loop (VCT_DEBUG_CONE_VARIANT 0, and the instrumented sample counter);
// cone_kernel_fast is the production march.  profiles/r01_ncu_s7.md, r01_cone_experiments_s7.txt: what binds it.
// FMA contraction is allowed here: the frame is compared against the oracle with a tolerance
// (max abs 2/255, PSNR >= 45 dB), not bit for bit.
#include <cuda_fp16.h>
#include "vct_internal.cuh"

namespace vct {

struct F3 { float x, y, z; };
__device__ __forceinline__ F3 f3(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ F3 operator+(F3 a, F3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ F3 operator-(F3 a, F3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ F3 operator*(F3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ F3 operator*(F3 a, F3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ F3 operator-(F3 a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ F3 cross(F3 a, F3 b) { return f3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
__device__ __forceinline__ float length(F3 a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ F3 normalize(F3 a) { float l = length(a); return f3(a.x / l, a.y / l, a.z / l); }
__device__ __forceinline__ F3 mix(F3 a, F3 b, float t) { return a * (1.0f - t) + b * t; }
// production march only: a / |a| through the hardware rsqrt (2 ulp) instead of sqrt + three IEEE divisions (~40 instructions).  A zero
// component stays exactly zero and a zero vector still gives NaN, so direction signs and the zero-weight tests are unchanged.
__device__ __forceinline__ F3 normalize_fast(F3 a) { return a * rsqrtf(dot(a, a)); }
__device__ __forceinline__ float lg2_fast(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }
__device__ __forceinline__ F3 reflect(F3 I, F3 N) { return I - N * (2.0f * dot(N, I)); }
__device__ __forceinline__ F3 refract(F3 I, F3 N, float eta) {
  float d = dot(N, I);
  float k = 1.0f - eta * eta * (1.0f - d * d);
  if (k < 0.0f) return f3(0.f, 0.f, 0.f);
  return I * eta - N * (eta * d + sqrtf(k));
}

struct TraceArgs {
  GridView grid;
  const float* world_pos;
  const float* normal;
  const uint32_t* material;
  uint32_t* frame;
  int W, H;
  const vct_material_t* mats;
  Lights lights;
  float cube_size;
  float cam_pos[3];
  vct_trace_params_t prm;
  unsigned long long* counts;  // [4] diffuse, shadow, specular, refraction (+[4] shaded pixels)
  int n_diffuse, n_slots;
  int n_jobs;                  // work items per live tile of cone_kernel_fast
  uint32_t* work_counter;      // next (job, tile) item of the persistent cone kernel; zeroed by tile_list_kernel
  int tile_k;                  // multi-GPU split of the frame: lattice step of screen_tile_owner for prm.tile_nranks
  uint32_t first_reserved_sm;  // persistent cone kernel: CTAs placed on an SM with %smid >= this retire at once (SMs left to another frame's front half)
  int grouped;                 // cone_out layout: 0 = [slot][pixel]; 1 = [job][pixel] with job 0 = SUM of the diffuse cones, 1 specular, 2 refraction, 3 + i shadow of light i
  const uint32_t* tile_list;   // live 8x4 tiles (tile_y * tiles_x + tile_x), built by tile_list_kernel
  uint32_t* tile_count;
  float4* cone_out;            // [slot][pixel] cone results (rgba)
  size_t npix;
  PeerView pv;                 // multi-GPU: where the finished pixels of this rank's tiles go (nranks <= 1: nowhere)
};

// byte k of a packed RGBA8 word as float (exact), without an I2F conversion
__device__ __forceinline__ float byte_f(uint32_t w, int k) {
  return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650 + k)) - 8388608.0f;
}
__device__ __forceinline__ void madd4(float acc[4], float w, uint32_t word) {
  acc[0] = fmaf(w, byte_f(word, 0), acc[0]);
  acc[1] = fmaf(w, byte_f(word, 1), acc[1]);
  acc[2] = fmaf(w, byte_f(word, 2), acc[2]);
  acc[3] = fmaf(w, byte_f(word, 3), acc[3]);
}

// true when the 2x2x2 filter footprint of `pos` at `level` holds no non-zero texel in any direction
// (dilated occupancy bits written by the mip stage) or lies wholly outside the grid: the level's
// contribution to the sample is then exactly zero in the reference as well.
__device__ __forceinline__ bool footprint_empty(const GridView& g, int level, F3 pos) {
  const int N = g.R >> level;
  const float fN = (float)N;
  const float ux = fmaf(pos.x, fN, -0.5f), uy = fmaf(pos.y, fN, -0.5f), uz = fmaf(pos.z, fN, -0.5f);
  if (!(ux > -1.0f && ux < fN && uy > -1.0f && uy < fN && uz > -1.0f && uz < fN)) return true;  // also NaN
  const int x = (int)floorf(ux) + 1, y = (int)floorf(uy) + 1, z = (int)floorf(uz) + 1;           // in [0, N]
  const uint32_t w = __ldg(g.docc[level] + ((size_t)z * (N + 1) + y) * occ_wpr(N) + (x >> 5));
  return ((w >> (x & 31)) & 1u) == 0u;
}

// one mip level of sample_voxel (voxel_cone_tracing.frag:80-86): adds
//   weight * (|d.x| * tex[ix] + |d.y| * tex[iy] + |d.z| * tex[iz]) (pos)   in BYTE units (0..255)
__device__ __forceinline__ void fetch_level(const GridView& g, int level, F3 pos, F3 adir, int ix, int iy, int iz, float weight, float acc[4]) {
  const int N = g.R >> level;
  const float fN = (float)N;
  const float ux = fmaf(pos.x, fN, -0.5f), uy = fmaf(pos.y, fN, -0.5f), uz = fmaf(pos.z, fN, -0.5f);
  // footprint wholly outside (also catches NaN): border colour = 0
  if (!(ux > -1.0f && ux < fN && uy > -1.0f && uy < fN && uz > -1.0f && uz < fN)) return;
  const float fx = floorf(ux), fy = floorf(uy), fz = floorf(uz);
  const int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
  const float ax = ux - fx, ay = uy - fy, az = uz - fz;
  const float wxs[2] = {1.0f - ax, ax}, wys[2] = {1.0f - ay, ay}, wzs[2] = {1.0f - az, az};
  if (level == 0) {
    float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int dz = 0; dz < 2; dz++) {
      const int z = z0 + dz;
      if ((unsigned)z >= (unsigned)N) continue;
#pragma unroll
      for (int dy = 0; dy < 2; dy++) {
        const int y = y0 + dy;
        if ((unsigned)y >= (unsigned)N) continue;
        const float wyz = wys[dy] * wzs[dz];
        const uint32_t* row = g.base + ((size_t)z * N + y) * N;
#pragma unroll
        for (int dx = 0; dx < 2; dx++) {
          const int x = x0 + dx;
          if ((unsigned)x >= (unsigned)N) continue;
          const uint32_t w = __ldg(row + x);
          if (w) madd4(t, wxs[dx] * wyz, w);
        }
      }
    }
    const float s = weight * ((adir.x + adir.y) + adir.z);
#pragma unroll
    for (int k = 0; k < 4; k++) acc[k] = fmaf(s, t[k], acc[k]);
  } else {
    // levels >= 1: the raw texels of the three directional volumes out of the stacked array (point fetches, exact texel centres);
    // the weights stay fp32 (oracle rule R7)
    float tx[4] = {0.f, 0.f, 0.f, 0.f}, ty[4] = {0.f, 0.f, 0.f, 0.f}, tz[4] = {0.f, 0.f, 0.f, 0.f};
    const float inv_n = 1.0f / fN, inv_d = 1.0f / (float)(6 * g.pitch[level]), lod = (float)(level - 1);
    const int pitch = g.pitch[level];
#pragma unroll
    for (int dz = 0; dz < 2; dz++) {
      const int z = z0 + dz;
      if ((unsigned)z >= (unsigned)N) continue;
#pragma unroll
      for (int dy = 0; dy < 2; dy++) {
        const int y = y0 + dy;
        if ((unsigned)y >= (unsigned)N) continue;
        const float wyz = wys[dy] * wzs[dz];
        const float cy = ((float)y + 0.5f) * inv_n;
#pragma unroll
        for (int dx = 0; dx < 2; dx++) {
          const int x = x0 + dx;
          if ((unsigned)x >= (unsigned)N) continue;
          const float w = wxs[dx] * wyz;
          const float cx = ((float)x + 0.5f) * inv_n;
          const uchar4 ca = tex3DLod<uchar4>(g.tex_pt, cx, cy, ((float)(z + ix * pitch) + 0.5f) * inv_d, lod);
          const uchar4 cb = tex3DLod<uchar4>(g.tex_pt, cx, cy, ((float)(z + iy * pitch) + 0.5f) * inv_d, lod);
          const uchar4 cc = tex3DLod<uchar4>(g.tex_pt, cx, cy, ((float)(z + iz * pitch) + 0.5f) * inv_d, lod);
          const uint32_t a = *reinterpret_cast<const uint32_t*>(&ca), b = *reinterpret_cast<const uint32_t*>(&cb), c = *reinterpret_cast<const uint32_t*>(&cc);
          if (a) madd4(tx, w, a);
          if (b) madd4(ty, w, b);
          if (c) madd4(tz, w, c);
        }
      }
    }
    const float sx = weight * adir.x, sy = weight * adir.y, sz = weight * adir.z;
#pragma unroll
    for (int k = 0; k < 4; k++) acc[k] = fmaf(sx, tx[k], fmaf(sy, ty[k], fmaf(sz, tz[k], acc[k])));
  }
}

// The same for the RGBA16F storage variant (vct_grid_create_ex): texels are four halves in linear per-direction buffers (level 0: one buffer
// for all six directions), filtered in fp32.  Byte units like fetch_level: a colour of 1.0 counts 255.
__device__ __forceinline__ void fetch_level_f16(const GridView& g, int level, F3 pos, F3 adir, int ix, int iy, int iz, float weight, float acc[4]) {
  const int N = g.R >> level;
  const float fN = (float)N;
  const float ux = fmaf(pos.x, fN, -0.5f), uy = fmaf(pos.y, fN, -0.5f), uz = fmaf(pos.z, fN, -0.5f);
  if (!(ux > -1.0f && ux < fN && uy > -1.0f && uy < fN && uz > -1.0f && uz < fN)) return;
  const float fx = floorf(ux), fy = floorf(uy), fz = floorf(uz);
  const int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
  const float ax = ux - fx, ay = uy - fy, az = uz - fz;
  const float wxs[2] = {1.0f - ax, ax}, wys[2] = {1.0f - ay, ay}, wzs[2] = {1.0f - az, az};
  const unsigned long long* tx = g.f16[level * 6 + ix];
  const unsigned long long* ty = g.f16[level * 6 + iy];
  const unsigned long long* tz = g.f16[level * 6 + iz];
  const float sx = 255.0f * weight * adir.x, sy = 255.0f * weight * adir.y, sz = 255.0f * weight * adir.z;
#pragma unroll
  for (int dz = 0; dz < 2; dz++) {
    const int z = z0 + dz;
    if ((unsigned)z >= (unsigned)N) continue;
#pragma unroll
    for (int dy = 0; dy < 2; dy++) {
      const int y = y0 + dy;
      if ((unsigned)y >= (unsigned)N) continue;
      const float wyz = wys[dy] * wzs[dz];
#pragma unroll
      for (int dx = 0; dx < 2; dx++) {
        const int x = x0 + dx;
        if ((unsigned)x >= (unsigned)N) continue;
        const float w = wxs[dx] * wyz;
        const size_t idx = ((size_t)z * N + y) * N + x;
        const unsigned long long a = __ldg(tx + idx), b = level == 0 ? a : __ldg(ty + idx), c = level == 0 ? a : __ldg(tz + idx);
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const float va = __half2float(__ushort_as_half((unsigned short)(a >> (16 * k))));
          const float vb = __half2float(__ushort_as_half((unsigned short)(b >> (16 * k))));
          const float vc = __half2float(__ushort_as_half((unsigned short)(c >> (16 * k))));
          acc[k] = fmaf(w, fmaf(sx, va, fmaf(sy, vb, sz * vc)), acc[k]);
        }
      }
    }
  }
}

// trace_cone (voxel_cone_tracing.frag:88-119).  Returns the number of loop iterations the reference
// would execute when COUNT is set (no early exit in that build).
// the three directional textureLod fetches of sample_voxel through the texture units:
//   acc += weight255 * (|d.x| * tex[ix] + |d.y| * tex[iy] + |d.z| * tex[iz])(pos, tex_lod)      (byte units)
__device__ __forceinline__ void fetch_tex(const GridView& g, int ix, int iy, int iz, F3 pos, F3 adir, float tex_lod, float weight255, float acc[4]) {
  // a direction whose weight |d.a| is exactly 0 contributes exactly 0 (axis-aligned cones of the box walls): no fetch.
  // Direction d of the stacked array starts at normalised depth d/6 (GridView).
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  const float zz = pos.z * g.tex_zs;
  const float4 a = adir.x != 0.0f ? tex3DLod<float4>(g.tex_lin, pos.x, pos.y, zz + (float)ix * (1.0f / 6.0f), tex_lod) : zero;
  const float4 b = adir.y != 0.0f ? tex3DLod<float4>(g.tex_lin, pos.x, pos.y, zz + (float)iy * (1.0f / 6.0f), tex_lod) : zero;
  const float4 c = adir.z != 0.0f ? tex3DLod<float4>(g.tex_lin, pos.x, pos.y, zz + (float)iz * (1.0f / 6.0f), tex_lod) : zero;
  const float sx = weight255 * adir.x, sy = weight255 * adir.y, sz = weight255 * adir.z;
  acc[0] = fmaf(sx, a.x, fmaf(sy, b.x, fmaf(sz, c.x, acc[0])));
  acc[1] = fmaf(sx, a.y, fmaf(sy, b.y, fmaf(sz, c.y, acc[1])));
  acc[2] = fmaf(sx, a.z, fmaf(sy, b.z, fmaf(sz, c.z, acc[2])));
  acc[3] = fmaf(sx, a.w, fmaf(sy, b.w, fmaf(sz, c.w, acc[3])));
}

template <bool COUNT, bool TEX, bool F16 = false>
__device__ __forceinline__ uint32_t trace_cone(const GridView& g, F3 origin, F3 dir, float aperture, float max_dist, float out[4]) {
  dir = normalize(dir);
  const int ix = dir.x < 0.0f ? 0 : 1, iy = dir.y < 0.0f ? 2 : 3, iz = dir.z < 0.0f ? 4 : 5;
  const F3 adir = f3(fabsf(dir.x), fabsf(dir.y), fabsf(dir.z));
  const float cube_res = (float)g.R;
  const float voxel_size = 1.0f / cube_res;
  const float max_level = (float)(g.levels - 1);
  const float margin = 0.5f / (float)(g.R >> (g.levels - 1));  // half a texel of the coarsest level
  float acc[4] = {0.f, 0.f, 0.f, 0.f};  // byte units
  float dist = 3.0f * voxel_size;
  if (!(dir.x == dir.x && dir.y == dir.y && dir.z == dir.z)) {
    // A direction that is not a number (normalize() of a zero vector: refract() under total reflection, a vertex normal of length zero).
    // In the reference every weight |d| * textureLod(..) of the first sample is NaN, the accumulator turns NaN and `alpha < 1` ends the
    // loop (voxel_cone_tracing.frag:98-116): the cone returns NaN whenever its loop runs at all, and the pixel ends up black.
    const bool runs = dist < max_dist;
#pragma unroll
    for (int c = 0; c < 4; c++) out[c] = runs ? __int_as_float(0x7FC00000) : 0.0f;
    return runs ? 1u : 0u;
  }
  float diam = dist * aperture;
  F3 sp = f3(fmaf(dir.x, dist, origin.x), fmaf(dir.y, dist, origin.y), fmaf(dir.z, dist, origin.z));
  uint32_t iters = 0;
  while (acc[3] < 255.0f && dist < max_dist) {
    if (!COUNT) {
      // the cone has left the border-padded cube for good: every later sample is exactly zero
      if ((sp.x <= -margin && dir.x <= 0.f) || (sp.x >= 1.0f + margin && dir.x >= 0.f) || (sp.y <= -margin && dir.y <= 0.f) ||
          (sp.y >= 1.0f + margin && dir.y >= 0.f) || (sp.z <= -margin && dir.z <= 0.f) || (sp.z >= 1.0f + margin && dir.z >= 0.f))
        break;
    }
    float lod = fmaxf(log2f(diam * cube_res), 0.0f);
    lod = fminf(lod, max_level);
    const float fl = floorf(lod);
    const int l0 = (int)fl;
    const float f = lod - fl;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    // levels whose filter footprint is all zero are skipped (exactly zero in the reference too)
    const bool e0 = footprint_empty(g, l0, sp);
    const bool e1 = (f > 0.0f) ? footprint_empty(g, l0 + 1, sp) : true;
    if (TEX) {
      if (lod < 1.0f) {
        if (!e0) fetch_level(g, 0, sp, adir, ix, iy, iz, 1.0f - lod, s);          // level 0 in software (shared by the three directions)
        if (!e1) fetch_tex(g, ix, iy, iz, sp, adir, 0.0f, 255.0f * lod, s);         // level 1 = array level 0
      } else if (!(e0 && e1)) {
        fetch_tex(g, ix, iy, iz, sp, adir, lod - 1.0f, 255.0f, s);                 // trilinear + mip-linear in the texture unit
      }
    } else if (F16) {
      if (!e0) fetch_level_f16(g, l0, sp, adir, ix, iy, iz, 1.0f - f, s);
      if (!e1) fetch_level_f16(g, l0 + 1, sp, adir, ix, iy, iz, f, s);
    } else {
      if (!e0) fetch_level(g, l0, sp, adir, ix, iy, iz, 1.0f - f, s);
      if (!e1) fetch_level(g, l0 + 1, sp, adir, ix, iy, iz, f, s);
    }
    const float k = 1.0f - acc[3] * (1.0f / 255.0f);
#pragma unroll
    for (int c = 0; c < 4; c++) acc[c] = fmaf(k, s[c], acc[c]);
    const float step = fmaxf(diam * 0.5f, voxel_size);
    dist = dist + step;
    diam = dist * aperture;
    sp = f3(fmaf(dir.x, dist, origin.x), fmaf(dir.y, dist, origin.y), fmaf(dir.z, dist, origin.z));
    iters++;
  }
#pragma unroll
  for (int c = 0; c < 4; c++) out[c] = acc[c] * (1.0f / 255.0f);
  return iters;
}


// ---------------------------------------------------------------------------------------------
// The production march.  Same samples, same order and same arithmetic per sample as trace_cone above; what changes is the
// bookkeeping around the fetch (measured with ncu: the literal loop spent 45 % of its issue slots on log2f, two occupancy
// tests and the cube-exit test per step):
//  * ONE occupancy test per step.  The occupancy bit of a texel of level >= 1 is the OR of its 8 children (mip stage), and the
//    2x2x2 footprint of level l around a point lies inside the children of the footprint of level l+1 around the same point
//    (x0 = floor(u - 1/2) is in {2X0, 2X0+1, 2X0+2} for X0 = floor(u/2 - 1/2)), so "footprint of l+1 empty" implies
//    "footprint of l empty": when both levels are blended only the coarser one is tested.
//  * the distance at which the cone has left the border-padded cube for good is computed once and folded into the loop bound.
//  * lod through the hardware lg2 (2 ulp-class error on a value that only weights two neighbouring levels).
__device__ __forceinline__ bool footprint_empty_fast(const GridView& g, int level, F3 pos) {
  const uint4 t = g.occ_tab[level];   // {word offset, N + 1, words per row, float bits of N}
  const float fN = __uint_as_float(t.w);
  // bit index = low corner of the two-texel footprint + 1 = floor(pos*N - 1/2) + 1 = the texel BOUNDARY nearest to pos*N, taken with
  // the 1.5*2^23 trick (one FFMA, no float->int conversion).  At an exact tie either neighbour may come out: the texel that drops
  // out of the tested footprint is the one whose filter weight is 0.  Negative and NaN positions give huge unsigned values.
  const uint32_t x = __float_as_uint(fmaf(pos.x, fN, 12582912.0f)) - 0x4B400000u;
  const uint32_t y = __float_as_uint(fmaf(pos.y, fN, 12582912.0f)) - 0x4B400000u;
  const uint32_t z = __float_as_uint(fmaf(pos.z, fN, 12582912.0f)) - 0x4B400000u;
  if (max(max(x, y), z) >= t.y) return true;   // footprint wholly outside the grid
  const uint32_t w = __ldg(g.docc_all + (t.x + (z * t.y + y) * t.z + (x >> 5)));
  return ((w >> (x & 31)) & 1u) == 0u;
}

// The three directional fetches of the production march.  All six volumes live in one array (GridView), so the texture object
// is the same for every lane and every direction -- a uniform register, one predicated TEX per direction -- and the direction
// is picked by the z offset zo.* = d/6.  (With one object per direction the handle differs between lanes and the compiler wraps
// every TEX in a "waterfall" loop: R2UR + vote + branch, 13 instructions per fetch instead of 2; 13 % of the kernel's issue slots.)
// ONE = the nearest-mip object.
template <bool ONE>
__device__ __forceinline__ void fetch_tex_u(const GridView& g, F3 zo, F3 pos, F3 adir, float tex_lod, float weight255, float acc[4]) {
  const cudaTextureObject_t t = ONE ? g.tex_one : g.tex_lin;
  const float zz = pos.z * g.tex_zs;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a, c = a;
  // a direction whose weight |d.a| is exactly 0 contributes exactly 0: no fetch
  if (adir.x != 0.0f) a = tex3DLod<float4>(t, pos.x, pos.y, zz + zo.x, tex_lod);
  if (adir.y != 0.0f) b = tex3DLod<float4>(t, pos.x, pos.y, zz + zo.y, tex_lod);
  if (adir.z != 0.0f) c = tex3DLod<float4>(t, pos.x, pos.y, zz + zo.z, tex_lod);
  const float sx = weight255 * adir.x, sy = weight255 * adir.y, sz = weight255 * adir.z;
  acc[0] = fmaf(sx, a.x, fmaf(sy, b.x, fmaf(sz, c.x, acc[0])));
  acc[1] = fmaf(sx, a.y, fmaf(sy, b.y, fmaf(sz, c.y, acc[1])));
  acc[2] = fmaf(sx, a.z, fmaf(sy, b.z, fmaf(sz, c.z, acc[2])));
  acc[3] = fmaf(sx, a.w, fmaf(sy, b.w, fmaf(sz, c.w, acc[3])));
}

// SPLIT (texture-unit sampler only): a fetch that needs ONE level goes through the nearest-mip texture object (g.tex_one),
// which filter one level instead of two -- half the work of the binding unit (the TEX pipe does not shortcut a zero LOD fraction:
// measured ~4 quad-cycles per trilinear + mip-linear request whatever the fraction).  One level suffices when
//  * the LOD is integral (lod < 1: level 1 is the only array level involved; lod clamped at the last level), or
//  * the finer of the two blended levels has an empty footprint: its term is exactly zero, what remains is f * coarser level.
//    (costs a second occupancy lookup on the LSU pipe, which has slack, only when the coarser footprint is not empty).
template <bool TEX, bool SPLIT>
__device__ __forceinline__ void trace_cone_fast(const GridView& g, bool alive, F3 origin, F3 dir, float aperture, float max_dist, float out[4]) {
  dir = normalize_fast(dir);
  const int ix = dir.x < 0.0f ? 0 : 1, iy = dir.y < 0.0f ? 2 : 3, iz = dir.z < 0.0f ? 4 : 5;
  const F3 adir = f3(fabsf(dir.x), fabsf(dir.y), fabsf(dir.z));
  const float cube_res = (float)g.R;
  const float voxel_size = 1.0f / cube_res;
  const float max_level = (float)(g.levels - 1);
  // past one texel of the coarsest level outside [0,1] every footprint of every level is wholly outside the grid
  const float margin = 1.0f / (float)(g.R >> (g.levels - 1));
  float end = max_dist;
  {
    const float o[3] = {origin.x, origin.y, origin.z}, d[3] = {dir.x, dir.y, dir.z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
      float t_exit;
      if (d[k] < 0.0f) t_exit = __fdividef(o[k] + margin, -d[k]);            // approximate division: the bound carries a voxel of slack
      else if (d[k] > 0.0f) t_exit = __fdividef((1.0f + margin) - o[k], d[k]);
      else t_exit = (o[k] <= -margin || o[k] >= 1.0f + margin) ? 0.0f : max_dist;
      end = fminf(end, t_exit + voxel_size);   // + one voxel: rounding slack, the extra samples are exactly zero
    }
    if (!(d[0] == d[0] && d[1] == d[1] && d[2] == d[2]) || !alive) end = 0.0f;   // NaN direction: no march, the result is set behind the loop
  }
  float acc[4] = {0.f, 0.f, 0.f, 0.f};  // byte units
  const F3 zo = f3((float)ix * (1.0f / 6.0f), (float)iy * (1.0f / 6.0f), (float)iz * (1.0f / 6.0f));   // where the cone's three directions start in the stacked array
  float dist = 3.0f * voxel_size;
  float diam = dist * aperture;
  while (acc[3] < 255.0f && dist < end) {
    const F3 sp = f3(fmaf(dir.x, dist, origin.x), fmaf(dir.y, dist, origin.y), fmaf(dir.z, dist, origin.z));
    const float lod = fminf(fmaxf(lg2_fast(diam * cube_res), 0.0f), max_level);
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    if (TEX) {
      if (lod < 1.0f) {
        // level 0 in software (shared by the three directions) + level 1 = array level 0 through the texture unit
        const bool e1 = lod > 0.0f ? footprint_empty_fast(g, 1, sp) : true;
        const bool e0 = (lod > 0.0f && e1) ? true : footprint_empty_fast(g, 0, sp);
        if (!e0) fetch_level(g, 0, sp, adir, ix, iy, iz, 1.0f - lod, s);
        if (!e1) fetch_tex_u<SPLIT>(g, zo, sp, adir, 0.0f, 255.0f * lod, s);
      } else {
        const float fl = floorf(lod);
        const int l0 = (int)fl;
        const bool two = lod > fl;
        if (!footprint_empty_fast(g, two ? l0 + 1 : l0, sp)) {
          if (!SPLIT) {
            fetch_tex_u<false>(g, zo, sp, adir, lod - 1.0f, 255.0f, s);
          } else {
            // one level (nearest-mip objects, integral array level) unless both levels contribute
            bool one = !two;
            float tl = fl - 1.0f, w = 255.0f;
            if (two && footprint_empty_fast(g, l0, sp)) { one = true; tl = fl; w = 255.0f * (lod - fl); }
            if (one) fetch_tex_u<SPLIT>(g, zo, sp, adir, tl, w, s);
            else fetch_tex_u<false>(g, zo, sp, adir, lod - 1.0f, 255.0f, s);
          }
        }
      }
    } else {
      const float fl = floorf(lod);
      const int l0 = (int)fl;
      const float f = lod - fl;
      const bool e1 = f > 0.0f ? footprint_empty_fast(g, l0 + 1, sp) : true;
      const bool e0 = (f > 0.0f && e1) ? true : footprint_empty_fast(g, l0, sp);
      if (!e0) fetch_level(g, l0, sp, adir, ix, iy, iz, 1.0f - f, s);
      if (!e1) fetch_level(g, l0 + 1, sp, adir, ix, iy, iz, f, s);
    }
    const float k = 1.0f - acc[3] * (1.0f / 255.0f);
#pragma unroll
    for (int c = 0; c < 4; c++) acc[c] = fmaf(k, s[c], acc[c]);
    dist = dist + fmaxf(diam * 0.5f, voxel_size);
    diam = dist * aperture;
  }
  // A direction that is not a number (normalize() of a zero vector: refract() under total reflection, a vertex normal of length zero).
  // In the reference every weight |d| * textureLod(..) of the first sample is NaN, the accumulator turns NaN and `alpha < 1` ends the
  // loop (voxel_cone_tracing.frag:98-116): the cone returns NaN whenever its loop runs at all, and the pixel ends up black.
  if (alive && !(dir.x == dir.x && dir.y == dir.y && dir.z == dir.z) && 3.0f * voxel_size < max_dist) {
#pragma unroll
    for (int c = 0; c < 4; c++) acc[c] = __int_as_float(0x7FC00000);
  }
#pragma unroll
  for (int c = 0; c < 4; c++) out[c] = acc[c] * (1.0f / 255.0f);
}

__device__ __forceinline__ F3 tangent(F3 n) {
  F3 t1 = cross(n, f3(0.f, 0.f, 1.f)), t2 = cross(n, f3(0.f, 1.f, 0.f));
  return length(t1) > length(t2) ? normalize(t1) : normalize(t2);
}

__device__ __forceinline__ F3 tangent_fast(F3 n) {
  const F3 t1 = cross(n, f3(0.f, 0.f, 1.f)), t2 = cross(n, f3(0.f, 1.f, 0.f));
  return length(t1) > length(t2) ? normalize_fast(t1) : normalize_fast(t2);
}

__device__ __forceinline__ float specular_aperture(float shininess) {
  float rough = sqrtf(2.0f / (shininess + 2.0f));
  float a = tanf(1.57079f * rough);
  return fminf(fmaxf(a, 0.0174533f), 3.14159265f);
}

__device__ __forceinline__ uint32_t pack_rgba8(const float v[4]) {
  uint32_t r = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) r |= ((uint32_t)rintf(clamp01(v[k]) * 255.0f)) << (8 * k);
  return r;
}

constexpr float kTan22_5 = 0.55785173935f;
constexpr float kMaxDistance = 1.73205080757f;
constexpr uint32_t kBackground = 0xFF404026u;  // (0.15,0.25,0.25,1) -> (38,64,64,255), renderer.cpp:398
constexpr int kConeWarps = 4;                  // warps (= tiles) per cone_kernel CTA

// NON-REFERENCE VARIANT (BASELINE.json config 5): 16 diffuse cones = the normal + a ring of 5 at 30 degrees + a ring of
// 10 at 60 degrees (azimuth 72 k and 36 k + 18 degrees), direction = n * cos(t) + (o1 * cos(p) + o2 * sin(p)) * sin(t), aperture
// 2 tan(15 degrees), equal weights.  Coefficients {cos t, sin t cos p, sin t sin p} rounded to float, the same literals as the oracle's.
__device__ constexpr float kCone16[16][3] = {
    {1.0f, 0.0f, 0.0f},
    {0.8660253882408142f, 0.5f, 0.0f},
    {0.8660253882408142f, 0.15450850129127502f, 0.4755282700061798f},
    {0.8660253882408142f, -0.404508501291275f, 0.29389262199401855f},
    {0.8660253882408142f, -0.404508501291275f, -0.29389262199401855f},
    {0.8660253882408142f, 0.15450850129127502f, -0.4755282700061798f},
    {0.5f, 0.8236390948295593f, 0.2676165699958801f},
    {0.5f, 0.5090369582176208f, 0.7006292939186096f},
    {0.5f, 0.0f, 0.8660253882408142f},
    {0.5f, -0.5090369582176208f, 0.7006292939186096f},
    {0.5f, -0.8236390948295593f, 0.2676165699958801f},
    {0.5f, -0.8236390948295593f, -0.2676165699958801f},
    {0.5f, -0.5090369582176208f, -0.7006292939186096f},
    {0.5f, 0.0f, -0.8660253882408142f},
    {0.5f, 0.5090369582176208f, -0.7006292939186096f},
    {0.5f, 0.8236390948295593f, -0.2676165699958801f},
};
constexpr float kAperture16 = 0.5358983874320984f;

// direction of diffuse cone `slot` from the tangent frame: the nine cones of voxel_cone_tracing.frag:153-165 (and their first five),
// or the 16-cone variant
__device__ __forceinline__ F3 diffuse_dir(F3 normal, F3 o1, F3 o2, int slot, int n_cones = 9) {
  if (n_cones == 16) return (normal * kCone16[slot][0] + o1 * kCone16[slot][1]) + o2 * kCone16[slot][2];
  switch (slot) {
    case 0: return normal;
    case 1: return mix(normal, o1, 0.5f);
    case 2: return mix(normal, -o1, 0.5f);
    case 3: return mix(normal, o2, 0.5f);
    case 4: return mix(normal, -o2, 0.5f);
    case 5: return mix(normal, (o1 + o2) * 0.5f, 0.5f);
    case 6: return mix(normal, -((o1 + o2) * 0.5f), 0.5f);
    case 7: return mix(normal, (o1 - o2) * 0.5f, 0.5f);
    default: return mix(normal, -((o1 - o2) * 0.5f), 0.5f);
  }
}

struct Pixel {
  size_t pix;
  uint32_t mat_id;
  bool in_frame, live;
  F3 world, normal, pos;
};

// G-buffer fetch + the within_cube test of main() (voxel_cone_tracing.frag:248-251)
__device__ __forceinline__ Pixel load_pixel(const TraceArgs& a, int px, int py) {
  Pixel p;
  p.in_frame = px < a.W && py < a.H;
  p.pix = (size_t)py * a.W + px;
  p.mat_id = p.in_frame ? a.material[p.pix] : VCT_NO_TRIANGLE;
  p.live = p.mat_id != VCT_NO_TRIANGLE;
  p.world = f3(0.f, 0.f, 0.f); p.normal = f3(0.f, 0.f, 1.f); p.pos = f3(0.f, 0.f, 0.f);
  if (p.live) {
    p.world = f3(a.world_pos[p.pix * 3], a.world_pos[p.pix * 3 + 1], a.world_pos[p.pix * 3 + 2]);
    p.normal = f3(a.normal[p.pix * 3], a.normal[p.pix * 3 + 1], a.normal[p.pix * 3 + 2]);
    p.pos = f3(0.5f * (p.world.x / a.cube_size) + 0.5f, 0.5f * (p.world.y / a.cube_size) + 0.5f, 0.5f * (p.world.z / a.cube_size) + 0.5f);
    p.live = fabsf(p.pos.x) < 1.0f && fabsf(p.pos.y) < 1.0f && fabsf(p.pos.z) < 1.0f;
  }
  return p;
}

__device__ __forceinline__ bool tile_is_mine(const TraceArgs& a, int tile_x, int tile_y) {
  // multi-GPU split: 32x32 screen tiles are dealt to the ranks on a diagonal lattice (screen_tile_owner)
  if (a.prm.tile_nranks <= 1) return true;
  return screen_tile_owner(tile_x / 4, tile_y / 8, a.prm.tile_nranks, a.tile_k) == a.prm.tile_rank;
}

// Compacts the 8x4 tiles that contain at least one shaded pixel.  One warp takes kTilesPerWarp consecutive tiles: the G-buffer
// loads of all of them are in flight together (the kernel is a chain of dependent latencies: material -> position -> atomic ->
// store; with one tile per warp and an atomic per tile it took 21 us at 1080p), and the warp reserves its list entries with ONE
// atomicAdd.
constexpr int kTilesPerWarp = 4;
__global__ void __launch_bounds__(256)
tile_list_kernel(const TraceArgs a, uint32_t* __restrict__ tile_list, uint32_t* __restrict__ tile_count) {
  const int lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.work_counter = 0u;   // the cone kernel of this frame starts at item 0
  const int tiles_x = (a.W + 7) / 8, tiles_y = (a.H + 3) / 4, n_tiles = tiles_x * tiles_y;
  const int first = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kTilesPerWarp;
  if (first >= n_tiles) return;
  bool live[kTilesPerWarp];
#pragma unroll
  for (int k = 0; k < kTilesPerWarp; k++) {
    const int tile = first + k;
    live[k] = false;
    if (tile < n_tiles) {
      const int tile_x = tile % tiles_x, tile_y = tile / tiles_x;
      if (tile_is_mine(a, tile_x, tile_y)) live[k] = load_pixel(a, tile_x * 8 + (lane & 7), tile_y * 4 + (lane >> 3)).live;
    }
  }
  uint32_t mask = 0;
#pragma unroll
  for (int k = 0; k < kTilesPerWarp; k++) mask |= (__any_sync(0xffffffffu, live[k]) ? 1u : 0u) << k;
  if (mask == 0u) return;
  uint32_t base = 0;
  if (lane == 0) base = atomicAdd(tile_count, (uint32_t)__popc(mask));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (lane < kTilesPerWarp && ((mask >> lane) & 1u)) tile_list[base + __popc(mask & ((1u << lane) - 1u))] = (uint32_t)(first + lane);
}

template <bool COUNT, bool TEX, bool F16 = false>
__global__ void __launch_bounds__(32 * kConeWarps)
cone_kernel(const TraceArgs a) {
  const int lane = threadIdx.x & 31;
  const uint32_t t = blockIdx.x * kConeWarps + (threadIdx.x >> 5);
  if (t >= *a.tile_count) return;
  // long cones first: blockIdx.y = 0 is the last slot (shadow cones, up to ~5x more steps than diffuse)
  const int slot = a.n_slots - 1 - (int)blockIdx.y;
  const int nd = a.n_diffuse;
  const uint32_t tile = a.tile_list[t];
  const int tiles_x = (a.W + 7) / 8;
  const int tile_x = tile % tiles_x, tile_y = tile / tiles_x;
  const Pixel p = load_pixel(a, tile_x * 8 + (lane & 7), tile_y * 4 + (lane >> 3));
  const vct_material_t* m = p.live ? a.mats + p.mat_id : a.mats;
  const F3 cam = f3(a.cam_pos[0], a.cam_pos[1], a.cam_pos[2]);
  const F3 normal = p.normal, pos = p.pos;

  float r[4] = {0.f, 0.f, 0.f, 0.f};
  uint32_t iters = 0;
  int kind = -1;  // 0 diffuse, 1 shadow, 2 specular, 3 refraction
  if (p.live) {
    if (slot < nd) {
      if (a.prm.enable_diffuse) {
        const F3 o1 = normalize(tangent(normal));
        const F3 o2 = normalize(cross(o1, normal));
        const F3 d = diffuse_dir(normal, o1, o2, slot, nd);
        iters = trace_cone<COUNT, TEX, F16>(a.grid, pos, d, nd == 16 ? kAperture16 : kTan22_5, kMaxDistance, r);
        kind = 0;
      }
    } else if (slot == nd) {
      if (a.prm.enable_specular) {
        const F3 view_dir = normalize(p.world - cam);
        const F3 sd = normalize(reflect(-view_dir, normal));
        iters = trace_cone<COUNT, TEX, F16>(a.grid, pos, sd, specular_aperture(m->shininess), kMaxDistance, r);
        kind = 2;
      }
    } else if (slot == nd + 1) {
      const bool transmissive = m->illum == 4 || m->illum == 6 || m->illum == 7 || m->illum == 9;
      if (transmissive && a.prm.enable_specular) {
        const F3 view_dir = normalize(p.world - cam);
        const F3 rd = refract(view_dir, normal, 1.0f / m->ior);
        iters = trace_cone<COUNT, TEX, F16>(a.grid, pos, rd, specular_aperture(m->shininess), kMaxDistance, r);
        kind = 3;
      }
    } else {
      const int li = slot - (nd + 2);
      if (li < a.lights.n && a.prm.enable_direct && a.prm.enable_shadow) {
        const vct_point_light_t& L = a.lights.l[li];
        const F3 lp = f3(0.5f * (L.position[0] / a.cube_size) + 0.5f, 0.5f * (L.position[1] / a.cube_size) + 0.5f,
                         0.5f * (L.position[2] / a.cube_size) + 0.5f);
        F3 ld = lp - pos;
        const float d = length(ld);
        ld = f3(ld.x / d, ld.y / d, ld.z / d);
        iters = trace_cone<COUNT, TEX, F16>(a.grid, pos, ld, 0.1f, d, r);
        kind = 1;
      }
    }
    a.cone_out[(size_t)slot * a.npix + p.pix] = make_float4(r[0], r[1], r[2], r[3]);
  }
  if (COUNT) {
    uint32_t it = iters;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) it += __shfl_xor_sync(0xffffffffu, it, o);
    int kk = kind;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kk = max(kk, __shfl_xor_sync(0xffffffffu, kk, o));
    if (lane == 0 && kk >= 0) atomicAdd(&a.counts[kk], (unsigned long long)it);
    if (slot == 0) {
      const uint32_t ball = __ballot_sync(0xffffffffu, p.live);
      if (lane == 0) atomicAdd(&a.counts[4], (unsigned long long)__popc(ball));
    }
  }
}


// cone parameters of one (pixel, slot): direction, aperture, max distance; false = this slot traces nothing for the pixel
__device__ __forceinline__ bool cone_setup(const TraceArgs& a, const Pixel& p, int slot, F3& d, float& aperture, float& max_dist) {
  const int nd = a.n_diffuse;
  const vct_material_t* m = a.mats + p.mat_id;
  const F3 normal = p.normal;
  aperture = kTan22_5;
  max_dist = kMaxDistance;
  d = normal;
  if (slot < nd) {
    if (!a.prm.enable_diffuse) return false;
    const F3 o1 = tangent_fast(normal);   // (the reference normalises it twice: a no-op up to an ulp)
    const F3 o2 = normalize_fast(cross(o1, normal));
    d = diffuse_dir(normal, o1, o2, slot, nd);
    if (nd == 16) aperture = kAperture16;
    return true;
  }
  const F3 cam = f3(a.cam_pos[0], a.cam_pos[1], a.cam_pos[2]);
  if (slot == nd) {
    if (!a.prm.enable_specular) return false;
    const F3 view_dir = normalize_fast(p.world - cam);
    d = reflect(-view_dir, normal);   // trace_cone_fast normalises
    aperture = specular_aperture(m->shininess);
    return true;
  }
  if (slot == nd + 1) {
    const bool transmissive = m->illum == 4 || m->illum == 6 || m->illum == 7 || m->illum == 9;
    if (!(transmissive && a.prm.enable_specular)) return false;
    const F3 view_dir = normalize_fast(p.world - cam);
    d = refract(view_dir, normal, 1.0f / m->ior);
    aperture = specular_aperture(m->shininess);
    return true;
  }
  const int li = slot - (nd + 2);
  if (!(li < a.lights.n && a.prm.enable_direct && a.prm.enable_shadow)) return false;
  const vct_point_light_t& L = a.lights.l[li];
  const F3 lp = f3(0.5f * (L.position[0] / a.cube_size) + 0.5f, 0.5f * (L.position[1] / a.cube_size) + 0.5f,
                   0.5f * (L.position[2] / a.cube_size) + 0.5f);
  const F3 ld = lp - p.pos;
  const float dl = length(ld);
  d = ld;   // trace_cone_fast normalises
  aperture = 0.1f;
  max_dist = dl;
  return true;
}

// GROUP: one warp marches ALL diffuse cones of its tile one after the other and stores their sum (added in slot order, as
// trace_diffuse does): the G-buffer fetch and the tangent frame are paid once instead of nine times (12 % of the kernel's
// instructions) and the cone-result buffer shrinks from 12 to 4 float4 per pixel (142 -> 47 MB written here and read by shade_kernel
// at 1080p).  The long job (blockIdx.y = 0) is scheduled first.
// one work item of the production march: job `job` of live tile `t` (all 32 lanes of a warp)
template <bool TEX, bool SPLIT, bool GROUP>
__device__ __forceinline__ void cone_work_item(const TraceArgs& a, uint32_t t, int job, int lane) {
  const uint32_t tile = a.tile_list[t];
  const int tiles_x = (a.W + 7) / 8;
  const int tile_x = tile % tiles_x, tile_y = tile / tiles_x;
  // lane -> pixel of the 8x4 tile: four consecutive lanes (one TEX quad) are a 2x2 pixel block, not a 4x1 row: the pixels of a quad then
  // agree more often on which fetch a sample needs, and the texture unit works on whole quads
  const int lx = (lane & 1) | ((lane >> 1) & 6), ly = ((lane >> 1) & 1) | ((lane >> 3) & 2);
  const Pixel p = load_pixel(a, tile_x * 8 + lx, tile_y * 4 + ly);
  if (!p.live) return;
  float r[4];
  if (GROUP) {
    if (job == 0) {
      float sum[3] = {0.f, 0.f, 0.f};
      if (a.prm.enable_diffuse) {
        const F3 o1 = tangent_fast(p.normal);
        const F3 o2 = normalize_fast(cross(o1, p.normal));
#pragma unroll 1
        for (int i = 0; i < a.n_diffuse; i++) {
          trace_cone_fast<TEX, SPLIT>(a.grid, true, p.pos, diffuse_dir(p.normal, o1, o2, i, a.n_diffuse), a.n_diffuse == 16 ? kAperture16 : kTan22_5, kMaxDistance, r);
          sum[0] = sum[0] + r[0]; sum[1] = sum[1] + r[1]; sum[2] = sum[2] + r[2];
        }
      }
      a.cone_out[p.pix] = make_float4(sum[0], sum[1], sum[2], 0.f);
      return;
    }
    F3 d = f3(0.f, 0.f, 1.f);
    float aperture = kTan22_5, max_dist = 0.f;
    const bool on = cone_setup(a, p, a.n_diffuse + job - 1, d, aperture, max_dist);
    trace_cone_fast<TEX, SPLIT>(a.grid, on, p.pos, d, aperture, max_dist, r);
    a.cone_out[(size_t)job * a.npix + p.pix] = make_float4(r[0], r[1], r[2], r[3]);
  } else {
    // long cones first: job 0 is the last slot (shadow cones, up to ~5x more steps than diffuse)
    const int slot = a.n_slots - 1 - job;
    F3 d = f3(0.f, 0.f, 1.f);
    float aperture = kTan22_5, max_dist = 0.f;
    const bool on = cone_setup(a, p, slot, d, aperture, max_dist);
    trace_cone_fast<TEX, SPLIT>(a.grid, on, p.pos, d, aperture, max_dist, r);
    a.cone_out[(size_t)slot * a.npix + p.pix] = make_float4(r[0], r[1], r[2], r[3]);
  }
}

// GROUP: one warp marches ALL diffuse cones of its tile one after the other and stores their sum (added in slot order, as
// trace_diffuse does): the G-buffer fetch and the tangent frame are paid once instead of nine times (12 % of the kernel's
// instructions) and the cone-result buffer shrinks from 12 to 4 float4 per pixel (142 -> 47 MB written here and read by shade_kernel
// at 1080p).
// PERSISTENT kernel: MIN_CTAS CTAs per SM stay resident and their warps take (job, tile) items from a global counter, job-major so
// that the long jobs start first.  The number of live tiles is only known on the device: a grid sized for every tile of the frame
// launches mostly empty CTAs (63 % at 1080p on one GPU, 95 % on eight), and dynamic items leave no tail of half-empty CTAs.
template <bool TEX, bool SPLIT, int MIN_CTAS, bool GROUP>
__global__ void __launch_bounds__(32 * kConeWarps, MIN_CTAS)
cone_kernel_fast(const TraceArgs a) {
  {
    uint32_t smid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    if (smid >= a.first_reserved_sm) return;
  }
  const int lane = threadIdx.x & 31;
  const uint32_t n_live = *a.tile_count;
  const uint32_t total = n_live * (uint32_t)a.n_jobs;
  for (;;) {
    uint32_t item = 0;
    if (lane == 0) item = atomicAdd(a.work_counter, 1u);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= total) break;
    const uint32_t job = item / n_live;
    cone_work_item<TEX, SPLIT, GROUP>(a, item - job * n_live, (int)job, lane);
  }
}

// the same items with one warp per (tile of the frame, job) of a grid sized on the host (VCT_DEBUG_CONE_GRID)
template <bool TEX, bool SPLIT, int MIN_CTAS, bool GROUP>
__global__ void __launch_bounds__(32 * kConeWarps, MIN_CTAS)
cone_kernel_grid(const TraceArgs a) {
  const uint32_t t = blockIdx.x * kConeWarps + (threadIdx.x >> 5);
  if (t >= *a.tile_count) return;
  cone_work_item<TEX, SPLIT, GROUP>(a, t, (int)blockIdx.y, threadIdx.x & 31);
}

// main() (voxel_cone_tracing.frag:246-275) for one pixel
__device__ __forceinline__ void store_pixel(const TraceArgs& a, size_t pix, uint32_t rgba) { a.frame[pix] = rgba; }

__device__ void shade_pixel(const TraceArgs& a, int tile_x, int tile_y, int lane) {
  const Pixel p = load_pixel(a, tile_x * 8 + (lane & 7), tile_y * 4 + (lane >> 3));
  if (!p.in_frame) return;
  if (!p.live) { store_pixel(a, p.pix, kBackground); return; }
  const vct_material_t* m = a.mats + p.mat_id;
  const F3 normal = p.normal, pos = p.pos;
  const int nd = a.n_diffuse;
  float out[4];
  if (a.prm.view_voxel_dir < 7) {
    // textureLod(tex3D[view_voxel_dir], pos, view_voxel_lod), blended over the clear colour
    const int d = a.prm.view_voxel_dir;
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    if (d >= 0 && d <= 5) {
      float lod = fminf(fmaxf(a.prm.view_voxel_lod, 0.0f), (float)(a.grid.levels - 1));
      const float fl = floorf(lod), f = lod - fl;
      const int l0 = (int)fl;
      const F3 ad = f3(d < 2 ? 1.f : 0.f, (d == 2 || d == 3) ? 1.f : 0.f, d >= 4 ? 1.f : 0.f);  // unit weight on direction d only
      const int ix = d < 2 ? d : 0, iy = (d == 2 || d == 3) ? d : 2, iz = d >= 4 ? d : 4;
      fetch_level(a.grid, l0, pos, ad, ix, iy, iz, 1.0f - f, t);
      if (f > 0.0f) fetch_level(a.grid, l0 + 1, pos, ad, ix, iy, iz, f, t);
    }
    const float bg[4] = {0.15f, 0.25f, 0.25f, 1.0f};
    const float al = t[3] * (1.0f / 255.0f);
#pragma unroll
    for (int k = 0; k < 4; k++) out[k] = (t[k] * (1.0f / 255.0f)) * al + bg[k] * (1.0f - al);
    store_pixel(a, p.pix, pack_rgba8(out));
    return;
  }
  const F3 cam = f3(a.cam_pos[0], a.cam_pos[1], a.cam_pos[2]);
  const F3 view_dir = normalize(p.world - cam);
  const F3 kd = f3(m->diffuse[0], m->diffuse[1], m->diffuse[2]);
  const F3 ks = f3(m->specular[0], m->specular[1], m->specular[2]);
  F3 fdiff = f3(0.f, 0.f, 0.f), fdir = f3(0.f, 0.f, 0.f), fspec = f3(0.f, 0.f, 0.f);
  if (a.prm.enable_diffuse) {
    F3 s = f3(0.f, 0.f, 0.f);
    if (a.grouped) {
      const float4 c = a.cone_out[p.pix];   // the cone kernel already added the nd cones, in the same order
      s = f3(c.x, c.y, c.z);
    } else {
      for (int i = 0; i < nd; i++) {
        const float4 c = a.cone_out[(size_t)i * a.npix + p.pix];
        s = s + f3(c.x, c.y, c.z);
      }
    }
    fdiff = kd * (s * (1.0f / (float)nd));
  }
  if (a.prm.enable_direct) {  // direct_light(), voxel_cone_tracing.frag:175-218
    F3 result = f3(0.f, 0.f, 0.f);
    for (int i = 0; i < a.lights.n; i++) {
      const vct_point_light_t& L = a.lights.l[i];
      const F3 lp = f3(0.5f * (L.position[0] / a.cube_size) + 0.5f, 0.5f * (L.position[1] / a.cube_size) + 0.5f,
                       0.5f * (L.position[2] / a.cube_size) + 0.5f);
      F3 ld = lp - pos;
      const float d = length(ld);
      ld = f3(ld.x / d, ld.y / d, ld.z / d);
      const float cos_surf = fmaxf(dot(normal, ld), 0.0f);
      const float att = 1.0f / (1.0f + d * d);
      const F3 light_color = f3(L.color[0], L.color[1], L.color[2]) * (att * cos_surf) * L.intensity;
      float shadow_level = 1.0f;
      if (a.prm.enable_shadow) shadow_level = fmaxf(0.0f, 1.0f - a.cone_out[(size_t)((a.grouped ? 3 : nd + 2) + i) * a.npix + p.pix].w);
      const float lambertian = fmaxf(dot(ld, normal), 0.0f);
      float refract_angle = 0.0f;
      if (m->dissolve <= 0.1f) {
        const F3 rf = refract(view_dir, normal, 1.0f / m->ior);
        refract_angle = fmaxf((1.0f - m->dissolve) * dot(rf, ld), 0.0f);
      }
      const F3 half_vec = normalize(ld + view_dir);
      float specular_angle = fmaxf(dot(half_vec, normal), 0.0f);
      specular_angle = fmaxf(specular_angle, refract_angle);
      const float specular_coeff = powf(specular_angle, m->shininess);
      const F3 brdf = kd * lambertian + ks * specular_coeff;
      result = result + (brdf * (shadow_level + 0.04f)) * light_color;
    }
    fdir = result + f3(clamp01(m->emission[0]), clamp01(m->emission[1]), clamp01(m->emission[2]));
  }
  if (a.prm.enable_specular) {
    const float4 c = a.cone_out[(size_t)(a.grouped ? 1 : nd) * a.npix + p.pix];
    fspec = ks * f3(c.x, c.y, c.z);
  }
  F3 rgb = (fspec + fdiff) + fdir;
  const bool transmissive = m->illum == 4 || m->illum == 6 || m->illum == 7 || m->illum == 9;
  if (transmissive && a.prm.enable_specular) {
    const float4 c = a.cone_out[(size_t)(a.grouped ? 2 : nd + 1) * a.npix + p.pix];
    const F3 rr = f3(m->transmittance[0], m->transmittance[1], m->transmittance[2]) * f3(c.x, c.y, c.z);
    rgb = mix(rr, rgb, m->dissolve);
  }
  out[0] = rgb.x; out[1] = rgb.y; out[2] = rgb.z; out[3] = 1.0f;
  store_pixel(a, p.pix, pack_rgba8(out));
}


__global__ void __launch_bounds__(256)
shade_kernel(const TraceArgs a) {
  const int tiles_x = (a.W + 7) / 8;
  const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int tile_x = tile % tiles_x, tile_y = tile / tiles_x;
  const bool mine = tile_y * 4 < a.H && tile_is_mine(a, tile_x, tile_y);
  if (mine) shade_pixel(a, tile_x, tile_y, lane);
}

// multi-GPU: the CTAs stride over this rank's 32x32 screen tiles and copy the finished pixels into the frame of the root rank
// (or of every rank) over NVLink with 16-byte stores (a tile row = 128 contiguous bytes); each CTA fences once, the last one publishes
// this rank's "tiles done" flag to the destination(s).  (Fusing the copy into shade_kernel was measured twice this round -- pixels pushed
// as they are shaded, and per 32x8 strip -- and lost: a system-scope fence per shading CTA costs more than this second pass, N = 2:
// shade + push 59 us here, 74 / 105 us fused.)
__global__ void __launch_bounds__(256)
frame_push_kernel(const TraceArgs a) {
  const int tiles_x = (a.W + 31) >> 5, tiles_y = (a.H + 31) >> 5;
  for (int tile = blockIdx.x; tile < tiles_x * tiles_y; tile += gridDim.x) {
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    if (screen_tile_owner(tx, ty, a.pv.nranks, a.tile_k) != a.pv.rank) continue;
    const int x0 = tx * 32, y0 = ty * 32;
    for (int u = threadIdx.x; u < 32 * 8; u += blockDim.x) {   // 32 rows x 8 uint4
      const int x = x0 + 4 * (u & 7), y = y0 + (u >> 3);
      if (y >= a.H || x >= a.W) continue;
      const size_t pix = (size_t)y * a.W + x;
      if (x + 3 < a.W && (pix & 3) == 0) {
        const uint4 v = *reinterpret_cast<const uint4*>(a.frame + pix);
        for (int p = 0; p < a.pv.nranks; p++)
          if (p != a.pv.rank && (a.pv.frame_root < 0 || p == a.pv.frame_root)) *reinterpret_cast<uint4*>(a.pv.frame[p] + pix) = v;
      } else {
        for (int c = 0; c < 4 && x + c < a.W; c++)
          for (int p = 0; p < a.pv.nranks; p++)
            if (p != a.pv.rank && (a.pv.frame_root < 0 || p == a.pv.frame_root)) a.pv.frame[p][pix + c] = a.frame[pix + c];
      }
    }
  }
  peer_signal_last_block(a.pv, PEER_FLAG_FRAME, a.pv.frame_root);
}

int launch_cone_trace(vct_device* dev, vct_scene* sc, vct_grid* g, const float* view, const vct_trace_params_t* p, vct_target_t_* t,
                      bool count_samples, const PeerView* push, int phase) {
  TraceArgs a;
  memset(&a.pv, 0, sizeof a.pv);
  if (push) a.pv = *push;
  a.grid = g->view();
  a.world_pos = t->world_pos; a.normal = t->normal; a.material = t->material; a.frame = t->frame;
  a.W = t->W; a.H = t->H;
  a.mats = sc->mats;
  a.lights = sc->lights;
  a.cube_size = sc->cube_size;
  // "camera_position" = column 3 of the VIEW matrix (src/renderer.cpp:279-280) -- reproduced, not fixed
  a.cam_pos[0] = view[12]; a.cam_pos[1] = view[13]; a.cam_pos[2] = view[14];
  a.prm = *p;
  if (a.prm.tile_nranks < 1) { a.prm.tile_nranks = 1; a.prm.tile_rank = 0; }
  a.tile_k = screen_tile_k(a.prm.tile_nranks);
  a.counts = (unsigned long long*)(dev->counters + 16);
  a.n_diffuse = p->n_diffuse_cones == 5 ? 5 : (p->n_diffuse_cones == 16 ? 16 : 9);
  a.n_slots = a.n_diffuse + 2 + sc->lights.n;
  a.grouped = 0;
  a.n_jobs = a.n_slots;
  a.work_counter = dev->counters + CNT_CONE_WORK;
  a.npix = (size_t)t->W * t->H;
  const int n_tiles = ((t->W + 7) / 8) * ((t->H + 3) / 4);
  // which march: 3 = production (one-level fetches through the nearest-mip texture object, all diffuse cones of a tile in one warp), 2 = one
  // warp per cone slot, 1 = every fetch blends two levels, 0 = literal loop; vct_debug_set(VCT_DEBUG_CONE_VARIANT) lets tests compare them
  const bool tex = p->sampler == VCT_SAMPLER_TEX && g->levels >= 2 && g->fmt == VCT_GRID_RGBA8;
  const bool ev = dev->debug_cone_variant >= 0;
  int variant = ev ? dev->debug_cone_variant : 3;
  // grouping makes the diffuse warps nine times longer: with few tiles per GPU (small frames, many ranks) the tail of the launch costs
  // more than the shared set-up saves (512x512: 200 -> 224 us; 1920x1080: 887 -> 863 us)
  // (and the fp32 software sampler, 72 registers, spills in the grouped form: 3.7 -> 4.6 ms at 1080p)
  if (!ev && (n_tiles / a.prm.tile_nranks < 32768 || !tex)) variant = 2;
  a.grouped = (!count_samples && variant >= 3) ? 1 : 0;
  // cone result buffer [slot or job][pixel] and the live-tile list, grown on demand
  const size_t need = (size_t)(a.grouped ? 3 + sc->lights.n : a.n_slots) * a.npix;
  if (need > t->cone_out_elems) {
    VCT_CUDA(cudaStreamSynchronize(dev->stream));
    cudaFree(t->cone_out);
    t->cone_out = nullptr; t->cone_out_elems = 0;
    VCT_CUDA(cudaMalloc(&t->cone_out, need * sizeof(float4)));
    t->cone_out_elems = need;
  }
  if (!t->tile_list) {
    VCT_CUDA(cudaMalloc(&t->tile_list, ((size_t)n_tiles + 1) * sizeof(uint32_t)));
  }
  a.tile_list = t->tile_list + 1;
  a.tile_count = t->tile_list;
  a.cone_out = (float4*)t->cone_out;
  cudaStream_t s = dev->stream;
  const bool debug_view = p->view_voxel_dir < 7;
  VCT_REQUIRE(!(debug_view && g->fmt == VCT_GRID_RGBA16F), "the voxel debug view reads RGBA8 grids only");
  if (!debug_view && phase != 2) {
    { int rc = launch_fill_u32(s, t->tile_list, 1, 0u); if (rc) return rc; }
    tile_list_kernel<<<(n_tiles + 8 * kTilesPerWarp - 1) / (8 * kTilesPerWarp), 256, 0, s>>>(a, t->tile_list + 1, t->tile_list);
  }
  if (phase == 1) {
    VCT_CUDA(cudaGetLastError());
    return VCT_OK;
  }
  if (!debug_view) {
    const dim3 grid((n_tiles + kConeWarps - 1) / kConeWarps, a.n_slots);
    const dim3 grid_jobs(grid.x, 3 + sc->lights.n);   // GROUP: the diffuse cones are one job
    VCT_CUDA(cudaEventRecord(dev->ev[6], s));
    if (g->fmt == VCT_GRID_RGBA16F) {   // storage variant: the literal march with fp32 filtering of the half texels
      if (count_samples) {
        VCT_CUDA(cudaMemsetAsync(dev->counters + 16, 0, 8 * sizeof(unsigned long long), s));
        cone_kernel<true, false, true><<<grid, 32 * kConeWarps, 0, s>>>(a);
      } else {
        cone_kernel<false, false, true><<<grid, 32 * kConeWarps, 0, s>>>(a);
      }
    } else if (count_samples) {
      VCT_CUDA(cudaMemsetAsync(dev->counters + 16, 0, 8 * sizeof(unsigned long long), s));
      cone_kernel<true, false><<<grid, 32 * kConeWarps, 0, s>>>(a);
    } else if (variant == 0) {
      if (tex) cone_kernel<false, true><<<grid, 32 * kConeWarps, 0, s>>>(a);
      else cone_kernel<false, false><<<grid, 32 * kConeWarps, 0, s>>>(a);
    } else {
      const bool persist = !dev->debug_cone_grid;
      const int sms = dev->prop.multiProcessorCount;
      a.first_reserved_sm = (uint32_t)(sms - dev->cone_reserved_sms);
      a.n_jobs = a.grouped ? (int)grid_jobs.y : a.n_slots;
      // launch K<TEX, SPLIT, MIN_CTAS, GROUP>: persistent (MIN_CTAS CTAs per SM) or one warp per (tile, job) of the whole frame
#define VCT_LAUNCH_CONE(TEXV, SPLITV, MINC, GROUPV)                                                                   \
  do {                                                                                                                \
    if (persist) cone_kernel_fast<TEXV, SPLITV, MINC, GROUPV><<<sms * (dev->cone_ctas_per_sm > 0 && dev->cone_ctas_per_sm < MINC ? dev->cone_ctas_per_sm : MINC), 32 * kConeWarps, 0, s>>>(a); \
    else cone_kernel_grid<TEXV, SPLITV, MINC, GROUPV><<<GROUPV ? grid_jobs : grid, 32 * kConeWarps, 0, s>>>(a);       \
  } while (0)
      if (variant == 1) {
        if (tex) VCT_LAUNCH_CONE(true, false, 9, false); else VCT_LAUNCH_CONE(false, false, 7, false);
      } else if (variant == 2) {
        if (tex) VCT_LAUNCH_CONE(true, true, 10, false); else VCT_LAUNCH_CONE(false, false, 7, false);
      } else {
        if (tex) VCT_LAUNCH_CONE(true, true, 10, true); else VCT_LAUNCH_CONE(false, false, 7, true);
      }
#undef VCT_LAUNCH_CONE
    }
    VCT_CUDA(cudaEventRecord(dev->ev[7], s));
  }
  shade_kernel<<<(n_tiles + 7) / 8, 256, 0, s>>>(a);
  if (a.pv.nranks > 1) {
    const int n32 = ((t->W + 31) / 32) * ((t->H + 31) / 32);
    frame_push_kernel<<<min(n32, dev->prop.multiProcessorCount * 4), 256, 0, s>>>(a);
  }
  VCT_CUDA(cudaGetLastError());
  return VCT_OK;
}

}  // namespace vct
