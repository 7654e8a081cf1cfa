// mip_arith.cuh -- the arithmetic of one anisotropic mip step (shader/mipmap.comp:22-100), shared by the mip
// kernels and -- compiled by g++ -- by the host-side arithmetic test (tests/test_mip_arith.py).
//
// The oracle (rules R5/R6) evaluates, per destination texel, direction d and channel k,
//     s = sum over the four (front, back) child pairs of  f_k + ((1 - f_a) * b_k)     (fp32, c/255.0f inputs, no FMA)
//     out = rint(clamp(s / 4, 0, 1) * 255)
// The same value in exact integers is  N / 1020  with
//     N = 255 * sum(F_k) + sum((255 - F_a) * B_k)   = two 4-way byte dot products (IDP.4A) once the eight children
// are transposed into per-channel byte vectors of the front and the back face (PRMT).  The fp32 chain of the
// oracle carries an error below 1e-4 in units of the result, and a non-tie N / 1020 is at least 1 / 1020 = 9.8e-4
// away from the nearest rounding boundary, so  rint(N / 1020)  IS the oracle's result unless N = 1020 m + 510
// exactly (a tie: which way the oracle goes then depends on its rounding errors).  Ties are detected exactly
// (residual |N - 1020 m| == 510) and replayed with the oracle's own fp32 recipe (mip_replay_channel).  Net effect:
// ~2.5x fewer FMA-pipe cycles per texel than the fp32 recipe (measured instruction mix in DESIGN.md 3.2) and
// bit-exact results.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define VCT_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#include <string.h>
#define VCT_HD inline
#endif

namespace vct {

// children numbering of mipmap.comp:10-20: bit 2 = (x == 0), bit 1 = (y == 0), bit 0 = (z == 0)
// => child i sits at offset (x,y,z) = (!(i>>2&1), !(i>>1&1), !(i&1)).
VCT_HD constexpr int child_id(int dx, int dy, int dz) { return ((dx ^ 1) << 2) | ((dy ^ 1) << 1) | (dz ^ 1); }

// ---- byte helpers (device: one instruction each) ----
VCT_HD uint32_t byte_perm(uint32_t x, uint32_t y, uint32_t s) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(x, y, s);
#else
  const uint64_t v = ((uint64_t)y << 32) | x;
  uint32_t r = 0;
  for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
  return r;
#endif
}
VCT_HD uint32_t dot4_u8(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
  return __dp4a(a, b, c);
#else
  for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 0xFFu) * ((b >> (8 * i)) & 0xFFu);
  return c;
#endif
}
VCT_HD float u32_as_f32(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}
VCT_HD uint32_t f32_as_u32(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
VCT_HD float fma_rn(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(a, b, c);
#else
  return fmaf(a, b, c);
#endif
}
VCT_HD float mul_rn(float a, float b) {   // a * b, never contracted into an FMA
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  volatile float r = a * b; return r;
#endif
}
VCT_HD float add_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b; return r;
#endif
}

// exact unorm8 -> float: fl(b / 255)  (multiply by fl(1/255) plus one exact Newton step; equal to the IEEE division for all 256 inputs)
VCT_HD float unorm8(uint32_t word, int byte) {
  const float b = u32_as_f32(byte_perm(word, 0x4B000000u, 0x7650u + (uint32_t)byte)) - 8388608.0f;   // byte -> float without I2F
  const float k = 0.003921568859368563f;  // fl(1/255)
  const float q = mul_rn(b, k);
  const float r = fma_rn(-q, 255.0f, b);  // exact residual
  return fma_rn(r, k, q);                 // correctly rounded quotient
}

// ---- the fp32 recipe of the oracle for ONE (direction, channel): used for ties only ----
// The (front, back) child pairs of mipmap.comp:59-98 IN THE SHADER'S ORDER (the fp32 sum is order dependent):
//   -x (0,4) (1,5) (2,6) (3,7)   -y (0,2) (1,3) (5,7) (4,6)   -z (0,1) (2,3) (4,5) (6,7);  +axis = the same pairs swapped.
// Packed one nibble per child: low half = the fronts of the negative direction, high half = its backs.
VCT_HD uint32_t mip_pair_word(int axis) { return axis == 0 ? 0x76543210u : (axis == 1 ? 0x67324510u : 0x75316420u); }
// load(i) returns child i (RGBA8 word).  Returns the destination byte of channel k in direction d.
template <class Load>
VCT_HD uint32_t mip_replay_channel(Load load, int d, int k) {
  const uint32_t word = mip_pair_word(d >> 1);
  const uint32_t fronts = (d & 1) ? word >> 16 : word & 0xFFFFu, backs = (d & 1) ? word & 0xFFFFu : word >> 16;
  float s = 0.f;
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const uint32_t f = load((int)((fronts >> (4 * p)) & 7u)), b = load((int)((backs >> (4 * p)) & 7u));
    const float v = add_rn(unorm8(f, k), mul_rn(add_rn(1.0f, -unorm8(f, 3)), unorm8(b, k)));   // alpha_blend, mipmap.comp:40-43
    s = p == 0 ? v : add_rn(s, v);
  }
  const float t = mul_rn(fminf(fmaxf(mul_rn(s, 0.25f), 0.0f), 1.0f), 255.0f);
  return f32_as_u32(add_rn(t, 12582912.0f)) & 0xFFu;   // rintf via 1.5 * 2^23 (ties to even, same as rintf)
}

// ---- per-channel byte vectors of the faces of the 2x2x2 child cube ----
// Stage 1 of a 4x4 byte transpose: two words (a, b) -> lo = [a.R, b.R, a.G, b.G], hi = [a.B, b.B, a.A, b.A]
struct BytePair { uint32_t lo, hi; };
VCT_HD BytePair pair_bytes(uint32_t a, uint32_t b) {
  BytePair p;
  p.lo = byte_perm(a, b, 0x5140u);
  p.hi = byte_perm(a, b, 0x7362u);
  return p;
}
// Stage 2: pairs (a, b) and (c, e) -> ch[k] = [a.k, b.k, c.k, e.k]
struct Face { uint32_t ch[4]; };
VCT_HD Face face_bytes(const BytePair& ab, const BytePair& ce) {
  Face f;
  f.ch[0] = byte_perm(ab.lo, ce.lo, 0x5410u);
  f.ch[1] = byte_perm(ab.lo, ce.lo, 0x7632u);
  f.ch[2] = byte_perm(ab.hi, ce.hi, 0x5410u);
  f.ch[3] = byte_perm(ab.hi, ce.hi, 0x7632u);
  return f;
}

// saturating pack of four non-negative ints into bytes: q0 | q1 << 8 | q2 << 16 | q3 << 24, each clamped to 255 (two I2IP instructions)
VCT_HD uint32_t pack4_sat_u8(uint32_t q0, uint32_t q1, uint32_t q2, uint32_t q3) {
#if defined(__CUDA_ARCH__)
  uint32_t hi, r;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(q3), "r"(q2), "r"(0u));
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(q1), "r"(q0), "r"(hi));
  return r;
#else
  const uint32_t c0 = q0 < 255u ? q0 : 255u, c1 = q1 < 255u ? q1 : 255u, c2 = q2 < 255u ? q2 : 255u, c3 = q3 < 255u ? q3 : 255u;
  return c0 | (c1 << 8) | (c2 << 16) | (c3 << 24);
#endif
}

// One direction: front face F, back face B (children in the same pair order) -> destination word, plus a 4-bit mask of
// the channels that are exact ties (to be replayed with mip_replay_channel).  N = 255 * sum(F_k) + sum((255 - F_a) * B_k);
// round(N / 1020) = floor(M / 1020) with M = N + 510 <= 520710.  One 32 x 32 -> 64 multiply gives both the quotient and the tie test:
// 4210753 * 1020 = 2^32 + 764, so for M = 1020 q + r the product is q * 2^32 + (764 q + r * 4210753) with the bracket < 2^32
// (<= 389640 + 1019 * 4210753): the high word IS q, and the low word is < 4210753 exactly when r = 0, i.e. on a tie.
// The clamp is the oracle's clamp(., 0, 1).
constexpr uint32_t kDiv1020 = 4210753u;
VCT_HD uint32_t mip_filter_faces(const Face& F, const Face& B, uint32_t& tie_mask) {
  const uint32_t W = ~F.ch[3];   // 255 - alpha of the four front children
  uint32_t q[4], lo[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const uint32_t M = dot4_u8(W, B.ch[k], dot4_u8(0xFFFFFFFFu, F.ch[k], 510u));
    const unsigned long long p = (unsigned long long)M * kDiv1020;
    q[k] = (uint32_t)(p >> 32);
    lo[k] = (uint32_t)p;
  }
  tie_mask = 0u;
  const uint32_t m01 = lo[0] < lo[1] ? lo[0] : lo[1], m23 = lo[2] < lo[3] ? lo[2] : lo[3];
  if ((m01 < m23 ? m01 : m23) < kDiv1020)   // rare: one branch per direction instead of a mask update per channel
    tie_mask = (lo[0] < kDiv1020 ? 1u : 0u) | (lo[1] < kDiv1020 ? 2u : 0u) | (lo[2] < kDiv1020 ? 4u : 0u) | (lo[3] < kDiv1020 ? 8u : 0u);
  return pack4_sat_u8(q[0], q[1], q[2], q[3]);
}

// All six directions of one destination texel whose eight children are the SAME words in every direction
// (level 0 -> 1: the reference writes the same value into all six level-0 textures, voxelize.frag:159-160).
// w[i] = child i.  out[d], and ties |= (channel mask) << (4 * d).
VCT_HD void mip_filter6_shared(const uint32_t (&w)[8], uint32_t (&out)[6], uint32_t& ties) {
  // x faces by a two-stage byte transpose of the z-pairs (0,1) (2,3) | (4,5) (6,7); the y and z faces are byte selections of the
  // two x faces (one PRMT per channel): y = 1: (0,1,4,5), y = 0: (2,3,6,7); z = 1: (0,2,4,6), z = 0: (1,3,5,7).  32 PRMT in all.
  const BytePair p01 = pair_bytes(w[0], w[1]), p23 = pair_bytes(w[2], w[3]), p45 = pair_bytes(w[4], w[5]), p67 = pair_bytes(w[6], w[7]);
  const Face x1 = face_bytes(p01, p23), x0 = face_bytes(p45, p67);   // pairs (0,4) (1,5) (2,6) (3,7)
  uint32_t tm;
  ties = 0u;
  out[0] = mip_filter_faces(x1, x0, tm); ties |= tm;
  out[1] = mip_filter_faces(x0, x1, tm); ties |= tm << 4;
  {
    Face y1, y0;   // pairs (0,2) (1,3) (4,6) (5,7)
#pragma unroll
    for (int k = 0; k < 4; k++) { y1.ch[k] = byte_perm(x1.ch[k], x0.ch[k], 0x5410u); y0.ch[k] = byte_perm(x1.ch[k], x0.ch[k], 0x7632u); }
    out[2] = mip_filter_faces(y1, y0, tm); ties |= tm << 8;
    out[3] = mip_filter_faces(y0, y1, tm); ties |= tm << 12;
  }
  {
    Face z1, z0;   // pairs (0,1) (2,3) (4,5) (6,7)
#pragma unroll
    for (int k = 0; k < 4; k++) { z1.ch[k] = byte_perm(x1.ch[k], x0.ch[k], 0x6420u); z0.ch[k] = byte_perm(x1.ch[k], x0.ch[k], 0x7531u); }
    out[4] = mip_filter_faces(z1, z0, tm); ties |= tm << 16;
    out[5] = mip_filter_faces(z0, z1, tm); ties |= tm << 20;
  }
}

// One direction of one destination texel from its own eight children (levels >= 1 -> next).  tie_mask: 4 channel bits.
VCT_HD uint32_t mip_filter1(const uint32_t (&w)[8], int d, uint32_t& tie_mask) {
  // front / back face of direction d in matching pair order
  uint32_t f0, f1, f2, f3, b0, b1, b2, b3;
  switch (d >> 1) {
    case 0: f0 = w[0]; f1 = w[1]; f2 = w[2]; f3 = w[3]; b0 = w[4]; b1 = w[5]; b2 = w[6]; b3 = w[7]; break;   // x
    case 1: f0 = w[0]; f1 = w[1]; f2 = w[4]; f3 = w[5]; b0 = w[2]; b1 = w[3]; b2 = w[6]; b3 = w[7]; break;   // y
    default: f0 = w[0]; f1 = w[2]; f2 = w[4]; f3 = w[6]; b0 = w[1]; b1 = w[3]; b2 = w[5]; b3 = w[7]; break;  // z
  }
  const Face hi = face_bytes(pair_bytes(f0, f1), pair_bytes(f2, f3));   // the face at coordinate 1 of the axis
  const Face lo = face_bytes(pair_bytes(b0, b1), pair_bytes(b2, b3));   // the face at coordinate 0
  return (d & 1) ? mip_filter_faces(lo, hi, tie_mask) : mip_filter_faces(hi, lo, tie_mask);
}

}  // namespace vct
