/*
 * vct_oracle.cpp -- CPU ORACLE: a plain C++ restatement of the reference's GLSL hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see vct_oracle.h).  PARITY: the shader arithmetic below is pinned, bit for bit, to the reference's
 * own GLSL text executed on the CPU (oracle/glsl_ref/, tests/test_glsl_ref.py); the fixed-function GL behaviour (rules R1-R3,
 * R6-R8, in vct_fixed_function.h) is held against Mesa llvmpipe running the reference's three passes (oracle/gl_ref/,
 * tests/test_gl_llvmpipe.py); R4 (fragment order) is a written rule -- GL defines no order.  Known-answer tests: tests/test_oracle_kat.py (derived by hand from
 * the shader text).
 *
 * What is restated (paths relative to /root/reference):
 *   shader/voxelize.vert:24-30, voxelize.geom:25-55, voxelize.frag:46-161
 *   shader/mipmap.comp:10-100 driven by src/renderer.cpp:283-314
 *   shader/voxel_cone_tracing.vert:22-28, voxel_cone_tracing.frag:56-275
 *   GL state of src/renderer.cpp:316-390, texture parameters of src/texture_3d.cpp:3-25
 *
 * Implementation-defined GL behaviour, FIXED HERE IN WRITING (the CUDA product follows the
 * same rules; checked against Mesa llvmpipe where that driver can run the pass, see vct_oracle.h):
 *   R1 viewport: xw = (x_ndc + 1) * (W * 0.5), yw likewise, zw = (z_ndc + 1) * 0.5; y up,
 *      row 0 = bottom.
 *   R2 coverage: vertices snapped to 1/256 pixel (rintf, ties-to-even); 64-bit integer edge
 *      functions evaluated at pixel centres (i+0.5, j+0.5); both windings; a centre exactly on
 *      an edge is covered iff the edge is a left edge (dy < 0 for the counter-clockwise
 *      orientation in y-up window space) or a top edge (dy == 0 && dx < 0); zero-area
 *      triangles produce nothing; x/y clipping = scissor to the viewport; triangles with a
 *      vertex beyond a +-2^21 pixel guard band are dropped.
 *   R2c camera pass: a triangle with a vertex in front of the near plane (z_c < -w_c) is clipped
 *      against that plane in clip space before the divide (see orc_gbuffer); the far plane is
 *      the per-fragment test of R3.
 *   R3 interpolation: b_k = float(E_k) / float(2A) from the SNAPPED positions; affine
 *      attributes ((b0*a0 + b1*a1) + b2*a2); perspective attributes
 *      ((q0*a0 + q1*a1) + q2*a2) / ((q0 + q1) + q2) with q_k = b_k * (1 / w_k);
 *      zw = (b0*zw0 + b1*zw1) + b2*zw2; near/far = per-fragment 0 <= zw <= 1.
 *   R4 fragment order for the order-dependent running average: draw order, triangle order in
 *      the index buffer, pixel row ascending, pixel column ascending (GL gives no guarantee).
 *   R5 float->int: truncation; GLSL round(): ties-to-even (rintf); normalize(v) = v / sqrt(dot);
 *      no fused multiply-add anywhere (build with -ffp-contract=off); left-to-right evaluation; min / max / clamp of a NaN
 *      (GLSL: undefined; a zero-length vertex normal produces one) return the non-NaN operand (fminf / fmaxf = IEEE minNum /
 *      maxNum = NVIDIA's FMNMX), so such a fragment stores colour 0.
 *   R6 unorm8 -> float: c / 255.0f; float -> unorm8: rintf(clamp(v,0,1) * 255.0f).
 *   R7 textureLod: lod clamped to [0, levels-1]; l0 = floor, l1 = min(l0+1, levels-1);
 *      per level u = s*N - 0.5, i0 = floor(u), fp32 trilinear weights, texels outside
 *      [0,N-1] = border (0,0,0,0) (CLAMP_TO_BORDER, default border); (1-f)*t0 + f*t1.
 *   R8 depth: GL_LESS on fp32 zw, first drawn wins ties; framebuffer RGBA8, cleared to
 *      (0.15,0.25,0.25,1) = (38,64,64,255).
 *   R9 normal matrix = mat3(transpose(inverse(model))) evaluated for an affine model matrix
 *      as cofactor(upper 3x3) / det in double precision, rounded to float.
 */
#include "vct_oracle.h"
#include "vct_fixed_function.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {
using namespace vct_ff;

inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 a, V3 b) {
  return v3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
inline float length(V3 a) { return sqrtf(dot(a, a)); }
inline V3 normalize(V3 a) { return divs(a, length(a)); }                 /* R5 */
inline V3 mix(V3 a, V3 b, float t) { return add(mul(a, 1.0f - t), mul(b, t)); }
inline V3 scale_and_bias(V3 p) { return v3(0.5f * p.x + 0.5f, 0.5f * p.y + 0.5f, 0.5f * p.z + 0.5f); }
inline V3 reflect(V3 I, V3 N) { return sub(I, mul(N, 2.0f * dot(N, I))); }
inline V3 refract(V3 I, V3 N, float eta) {
  float d = dot(N, I);
  float k = 1.0f - eta * eta * (1.0f - d * d);
  if (k < 0.0f) return v3(0, 0, 0);
  return sub(mul(I, eta), mul(N, eta * d + sqrtf(k)));
}

/* column-major mat4 * (x,y,z,1) -- GLSL M*v = ((c0*x + c1*y) + c2*z) + c3*w */
inline V4 mat4_mul_point(const float* m, V3 p) {
  V4 r;
  r.x = ((m[0] * p.x + m[4] * p.y) + m[8] * p.z) + m[12];
  r.y = ((m[1] * p.x + m[5] * p.y) + m[9] * p.z) + m[13];
  r.z = ((m[2] * p.x + m[6] * p.y) + m[10] * p.z) + m[14];
  r.w = ((m[3] * p.x + m[7] * p.y) + m[11] * p.z) + m[15];
  return r;
}
inline V4 mat4_mul_v4(const float* m, V4 p) {
  V4 r;
  r.x = ((m[0] * p.x + m[4] * p.y) + m[8] * p.z) + m[12] * p.w;
  r.y = ((m[1] * p.x + m[5] * p.y) + m[9] * p.z) + m[13] * p.w;
  r.z = ((m[2] * p.x + m[6] * p.y) + m[10] * p.z) + m[14] * p.w;
  r.w = ((m[3] * p.x + m[7] * p.y) + m[11] * p.z) + m[15] * p.w;
  return r;
}
void mat4_mul(const float* a, const float* b, float* out) { /* out = a*b, column-major */
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 4; r++)
      out[c * 4 + r] = ((a[0 * 4 + r] * b[c * 4 + 0] + a[1 * 4 + r] * b[c * 4 + 1]) + a[2 * 4 + r] * b[c * 4 + 2]) +
                       a[3 * 4 + r] * b[c * 4 + 3];
}

/* R9: normal matrix (3x3 column-major) of an affine model matrix */
void normal_matrix(const float* m, float* nm) {
  double a00 = m[0], a10 = m[1], a20 = m[2];   /* column 0 */
  double a01 = m[4], a11 = m[5], a21 = m[6];   /* column 1 */
  double a02 = m[8], a12 = m[9], a22 = m[10];  /* column 2 */
  double c00 = a11 * a22 - a12 * a21, c01 = a12 * a20 - a10 * a22, c02 = a10 * a21 - a11 * a20;
  double c10 = a02 * a21 - a01 * a22, c11 = a00 * a22 - a02 * a20, c12 = a01 * a20 - a00 * a21;
  double c20 = a01 * a12 - a02 * a11, c21 = a02 * a10 - a00 * a12, c22 = a00 * a11 - a01 * a10;
  double det = (a00 * c00 + a01 * c01) + a02 * c02;
  /* inverse-transpose = cofactor / det; element (row r, col c) = C_rc / det */
  nm[0] = (float)(c00 / det); nm[1] = (float)(c10 / det); nm[2] = (float)(c20 / det); /* column 0: rows 0..2 */
  nm[3] = (float)(c01 / det); nm[4] = (float)(c11 / det); nm[5] = (float)(c21 / det);
  nm[6] = (float)(c02 / det); nm[7] = (float)(c12 / det); nm[8] = (float)(c22 / det);
}
inline V3 mat3_mul(const float* m, V3 v) {
  return v3((m[0] * v.x + m[3] * v.y) + m[6] * v.z, (m[1] * v.x + m[4] * v.y) + m[7] * v.z,
            (m[2] * v.x + m[5] * v.y) + m[8] * v.z);
}

/* ------------------------------------------------------------------ */
/* V5: imageAtomicRGBA8Avg (voxelize.frag:66-120), one sequential step */
inline uint32_t conv_vec4_to_rgba8(const float v[4]) { /* :66-71 uint(val) & 0xFF */
  return (((uint32_t)v[3]) & 0xFFu) << 24 | (((uint32_t)v[2]) & 0xFFu) << 16 | (((uint32_t)v[1]) & 0xFFu) << 8 |
         (((uint32_t)v[0]) & 0xFFu);
}
inline uint32_t enc_nibble(uint32_t m, uint32_t n) { /* :80-86 */
  return (m & 0xFEFEFEFEu) | (n & 1u) | (n & 2u) << 7 | (n & 4u) << 14 | (n & 8u) << 21;
}
inline uint32_t dec_nibble(uint32_t m) { /* :88-93 */
  return (m & 1u) | (m & 0x100u) >> 7 | (m & 0x10000u) >> 14 | (m & 0x1000000u) >> 21;
}
/* val255 = val * 255 (voxelize.frag:99).  Serialised CAS loop: the first CAS (expected 0) succeeds iff
 * the stored word is 0; otherwise exactly one averaging step is applied (:106-119). */
inline uint32_t rgba8_avg_fold(uint32_t stored, const float val255[4]) {
  if (stored == 0u) return enc_nibble(conv_vec4_to_rgba8(val255), 1u);
  uint32_t c = stored & 0xFEFEFEFEu;
  float r[4] = {(float)(c & 0xFFu), (float)((c >> 8) & 0xFFu), (float)((c >> 16) & 0xFFu), (float)((c >> 24) & 0xFFu)};
  uint32_t n = dec_nibble(stored);
  float fn = (float)n;
  n = n + 1u;
  float fn1 = (float)n;
  for (int k = 0; k < 4; k++) {
    float t = r[k] * fn + val255[k];
    t = t / fn1;
    t = rintf(t / 2.0f) * 2.0f; /* R5 */
    r[k] = t;
  }
  return enc_nibble(conv_vec4_to_rgba8(r), n);
}

/* V2: voxelize.geom:25-55. 0 -> project to (x,y); 1 -> (y,z); 2 -> (x,z) */
inline int select_axis(V3 w0, V3 w1, V3 w2) {
  V3 p1 = sub(w1, w0), p2 = sub(w2, w0);
  V3 c = cross(p1, p2);
  float px = fabsf(c.x), py = fabsf(c.y), pz = fabsf(c.z);
  if (pz > px && pz > py) return 0;
  if (px > py && px > pz) return 1;
  return 2;
}


/* C2: voxel_cone_tracing.frag:71-86 */
inline void sample_voxel(const Pyramid& p, V3 pos, V3 dir, float lod, float out[4]) {
  int ix = dir.x < 0.0f ? 0 : 1, iy = dir.y < 0.0f ? 2 : 3, iz = dir.z < 0.0f ? 4 : 5;
  float ax = fabsf(dir.x), ay = fabsf(dir.y), az = fabsf(dir.z);
  float tx[4], ty[4], tz[4];
  texture_lod(p, ix, pos, lod, tx);
  texture_lod(p, iy, pos, lod, ty);
  texture_lod(p, iz, pos, lod, tz);
  for (int k = 0; k < 4; k++) out[k] = (ax * tx[k] + ay * ty[k]) + az * tz[k];
}

/* C3: voxel_cone_tracing.frag:88-119 */
inline int trace_cone(const Pyramid& p, V3 origin, V3 dir, float aperture, float max_dist, float out[4]) {
  dir = normalize(dir);
  float cube_res = (float)p.R;
  float voxel_size = 1.0f / cube_res;
  float acc[4] = {0, 0, 0, 0};
  float dist = 3.0f * voxel_size;
  float diam = dist * aperture;
  V3 sp = add(mul(dir, dist), origin);
  int iters = 0;
  while (acc[3] < 1.0f && dist < max_dist) {
    float mip = fmaxf(log2f(diam * cube_res), 0.0f);
    float s[4];
    sample_voxel(p, sp, dir, mip, s);
    float k = 1.0f - acc[3];
    for (int c = 0; c < 4; c++) acc[c] = acc[c] + k * s[c];
    float step = fmaxf(diam / 2.0f, voxel_size);
    dist = dist + step;
    diam = dist * aperture;
    sp = add(mul(dir, dist), origin);
    iters++;
  }
  for (int c = 0; c < 4; c++) out[c] = acc[c];
  return iters;
}

inline V3 tangent(V3 n) { /* :126-134 */
  V3 t1 = cross(n, v3(0, 0, 1)), t2 = cross(n, v3(0, 1, 0));
  if (length(t1) > length(t2)) return normalize(t1);
  return normalize(t2);
}

const float TAN_22_5 = 0.55785173935f;
const float MAX_DISTANCE = 1.73205080757f;

inline float specular_aperture(float shininess) { /* :220-227 */
  float rough = sqrtf(2.0f / (shininess + 2.0f));
  float a = tanf(1.57079f * rough);
  return fminf(fmaxf(a, 0.0174533f), 3.14159265f);
}

struct ShadeCtx {
  const orc_scene_t* scene;
  Pyramid pyr;
  const orc_trace_params_t* prm;
  V3 camera_position;
};

struct SampleCount { uint64_t diffuse = 0, shadow = 0, specular = 0, refraction = 0; };

// NON-REFERENCE VARIANT (BASELINE.json config 5): 16 diffuse cones = the normal + a ring of 5 at 30 degrees + a ring of
// 10 at 60 degrees (azimuth 72 k and 36 k + 18 degrees), direction = n * cos(t) + (o1 * cos(p) + o2 * sin(p)) * sin(t) with the
// reference's tangent frame, aperture 2 tan(15 degrees), equal weights.  Coefficients {cos t, sin t cos p, sin t sin p} rounded to float.
static const float kCone16[16][3] = {
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
static const float kAperture16 = 0.5358983874320984f;

inline V3 trace_diffuse(const ShadeCtx& c, V3 origin, V3 normal, SampleCount& sc) { /* :140-168 */
  const float angle_mix = 0.5f;
  V3 o1 = normalize(tangent(normal));
  V3 o2 = normalize(cross(o1, normal));
  if (c.prm->n_diffuse_cones == 16) {
    float acc[3] = {0, 0, 0};
    for (int i = 0; i < 16; i++) {
      V3 d = add(add(mul(normal, kCone16[i][0]), mul(o1, kCone16[i][1])), mul(o2, kCone16[i][2]));
      float r[4];
      sc.diffuse += (uint64_t)trace_cone(c.pyr, origin, d, kAperture16, MAX_DISTANCE, r);
      acc[0] = acc[0] + r[0]; acc[1] = acc[1] + r[1]; acc[2] = acc[2] + r[2];
    }
    return v3(acc[0] / 16.0f, acc[1] / 16.0f, acc[2] / 16.0f);
  }
  V3 c1 = mul(add(o1, o2), 0.5f);
  V3 c2 = mul(sub(o1, o2), 0.5f);
  V3 dirs[9] = {normal,
                mix(normal, o1, angle_mix), mix(normal, neg(o1), angle_mix),
                mix(normal, o2, angle_mix), mix(normal, neg(o2), angle_mix),
                mix(normal, c1, angle_mix), mix(normal, neg(c1), angle_mix),
                mix(normal, c2, angle_mix), mix(normal, neg(c2), angle_mix)};
  int n = c.prm->n_diffuse_cones == 5 ? 5 : 9;
  float acc[3] = {0, 0, 0};
  for (int i = 0; i < n; i++) {
    float r[4];
    sc.diffuse += (uint64_t)trace_cone(c.pyr, origin, dirs[i], TAN_22_5, MAX_DISTANCE, r);
    acc[0] = acc[0] + r[0]; acc[1] = acc[1] + r[1]; acc[2] = acc[2] + r[2];
  }
  float d = (float)n;
  return v3(acc[0] / d, acc[1] / d, acc[2] / d);
}

inline V3 direct_light(const ShadeCtx& c, const orc_material_t& m, V3 pos, V3 normal, V3 view_dir, SampleCount& sc) { /* :175-218 */
  V3 result = v3(0, 0, 0);
  uint32_t nl = c.scene->n_lights < 10u ? c.scene->n_lights : 10u;
  V3 diffuse = v3(m.diffuse[0], m.diffuse[1], m.diffuse[2]);
  V3 specular = v3(m.specular[0], m.specular[1], m.specular[2]);
  for (uint32_t i = 0; i < nl; i++) {
    const orc_light_t& L = c.scene->lights[i];
    V3 lp = scale_and_bias(divs(v3(L.position[0], L.position[1], L.position[2]), c.scene->cube_size));
    V3 ld = sub(lp, pos);
    float d = length(ld);
    ld = divs(ld, d);
    float cos_surf = fmaxf(dot(normal, ld), 0.0f);
    float att = 1.0f / ((1.0f + 0.0f * d) + (1.0f * d) * d);
    V3 light_color = mul(mul(v3(L.color[0], L.color[1], L.color[2]), att * cos_surf), L.intensity);
    float shadow_level = 1.0f;
    if (c.prm->enable_shadow) {
      float r[4];
      sc.shadow += (uint64_t)trace_cone(c.pyr, pos, ld, 0.1f, d, r);
      shadow_level = fmaxf(0.0f, 1.0f - r[3]);
    }
    float lambertian = fmaxf(dot(ld, normal), 0.0f);
    float refract_angle = 0.0f;
    if (m.dissolve <= 0.1f) {
      V3 rf = refract(view_dir, normal, 1.0f / m.ior);
      refract_angle = fmaxf((1.0f - m.dissolve) * dot(rf, ld), 0.0f);
    }
    V3 half_vec = normalize(add(ld, view_dir));
    float specular_angle = fmaxf(dot(half_vec, normal), 0.0f);
    specular_angle = fmaxf(specular_angle, refract_angle);
    float specular_coeff = powf(specular_angle, m.shininess);
    V3 brdf = add(mul(diffuse, lambertian), mul(specular, specular_coeff));
    result = add(result, mulv(mul(brdf, shadow_level + 0.04f), light_color));
  }
  return add(result, v3(clamp01(m.emission[0]), clamp01(m.emission[1]), clamp01(m.emission[2])));
}

/* C6 main(): returns false when the fragment returns before writing (outside the cube) */
inline bool shade(const ShadeCtx& c, const orc_material_t& m, V3 world, V3 normal, float out[4], SampleCount& sc) {
  V3 pos = scale_and_bias(divs(world, c.scene->cube_size));
  if (!(fabsf(pos.x) < 1.0f && fabsf(pos.y) < 1.0f && fabsf(pos.z) < 1.0f)) return false;
  if (c.prm->view_voxel_dir < 7) {
    int d = c.prm->view_voxel_dir;
    if (d < 0 || d > 5) { out[0] = out[1] = out[2] = out[3] = 0.0f; return true; }
    texture_lod(c.pyr, d, pos, c.prm->view_voxel_lod, out);
    return true;
  }
  V3 view_dir = normalize(sub(world, c.camera_position));
  V3 fd = v3(0, 0, 0), fdir = v3(0, 0, 0), fs = v3(0, 0, 0);
  float sap = specular_aperture(m.shininess);
  if (c.prm->enable_diffuse) fd = mulv(v3(m.diffuse[0], m.diffuse[1], m.diffuse[2]), trace_diffuse(c, pos, normal, sc));
  if (c.prm->enable_direct) fdir = direct_light(c, m, pos, normal, view_dir, sc);
  if (c.prm->enable_specular) {
    V3 sd = normalize(reflect(neg(view_dir), normal));
    float r[4];
    sc.specular += (uint64_t)trace_cone(c.pyr, pos, sd, sap, MAX_DISTANCE, r);
    fs = mulv(v3(m.specular[0], m.specular[1], m.specular[2]), v3(r[0], r[1], r[2]));
  }
  V3 rgb = add(add(fs, fd), fdir);
  bool transmissive = (m.illum == 4 || m.illum == 6 || m.illum == 7 || m.illum == 9);
  if (transmissive && c.prm->enable_specular) {
    V3 rd = refract(view_dir, normal, 1.0f / m.ior);
    float r[4];
    sc.refraction += (uint64_t)trace_cone(c.pyr, pos, rd, sap, MAX_DISTANCE, r);
    V3 rr = mulv(v3(m.transmittance[0], m.transmittance[1], m.transmittance[2]), v3(r[0], r[1], r[2]));
    rgb = mix(rr, rgb, m.dissolve);
  }
  out[0] = rgb.x; out[1] = rgb.y; out[2] = rgb.z; out[3] = 1.0f;
  return true;
}

}  // namespace

/* ================================================================== */
extern "C" {

uint32_t orc_rgba8_avg_fold(uint32_t stored, const float val01[4]) {
  float v[4] = {val01[0] * 255.0f, val01[1] * 255.0f, val01[2] * 255.0f, val01[3] * 255.0f};
  return rgba8_avg_fold(stored, v);
}

int orc_select_axis(const float a[3], const float b[3], const float c[3]) {
  return select_axis(v3(a[0], a[1], a[2]), v3(b[0], b[1], b[2]), v3(c[0], c[1], c[2]));
}

void orc_perspective(float fovy, float aspect, float zn, float zf, float out[16]) {
  /* glm 0.9.9 perspectiveRH_NO (thirdparty/glm/glm/gtc/matrix_transform.inl:343-356); fovy in radians */
  float t = tanf(fovy / 2.0f);
  for (int i = 0; i < 16; i++) out[i] = 0.0f;
  out[0] = 1.0f / (aspect * t);
  out[5] = 1.0f / t;
  out[10] = -(zf + zn) / (zf - zn);
  out[11] = -1.0f;
  out[14] = -(2.0f * zf * zn) / (zf - zn);
}

void orc_look_at(const float eye[3], const float center[3], const float up[3], float out[16]) {
  /* glm lookAtRH */
  V3 e = v3(eye[0], eye[1], eye[2]);
  V3 f = normalize(sub(v3(center[0], center[1], center[2]), e));
  V3 s = normalize(cross(f, v3(up[0], up[1], up[2])));
  V3 u = cross(s, f);
  out[0] = s.x; out[4] = s.y; out[8] = s.z;
  out[1] = u.x; out[5] = u.y; out[9] = u.z;
  out[2] = -f.x; out[6] = -f.y; out[10] = -f.z;
  out[3] = 0; out[7] = 0; out[11] = 0;
  out[12] = -dot(s, e); out[13] = -dot(u, e); out[14] = dot(f, e); out[15] = 1.0f;
}

void orc_camera_front(float pitch_deg, float yaw_deg, float out[3]) {
  /* src/camera.h:25-37 */
  const float k = 0.01745329251994329576923690768489f;
  float p = pitch_deg * k, y = yaw_deg * k;
  V3 f = v3(cosf(p) * cosf(y), sinf(p), cosf(p) * sinf(y));
  f = normalize(f);
  out[0] = f.x; out[1] = f.y; out[2] = f.z;
}

float orc_specular_aperture(float shininess) { return specular_aperture(shininess); }

int orc_trace_cone_fmt(const uint32_t* const* levels, int R, int n_levels, int fmt, const float origin[3], const float dir[3],
                       float aperture, float max_dist, float out_rgba[4]) {
  Pyramid p{levels, R, n_levels, fmt};
  return trace_cone(p, v3(origin[0], origin[1], origin[2]), v3(dir[0], dir[1], dir[2]), aperture, max_dist, out_rgba);
}
int orc_trace_cone(const uint32_t* const* levels, int R, int n_levels, const float origin[3], const float dir[3],
                   float aperture, float max_dist, float out_rgba[4]) {
  return orc_trace_cone_fmt(levels, R, n_levels, 0, origin, dir, aperture, max_dist, out_rgba);
}

void orc_texture_lod(const uint32_t* const* levels, int R, int n_levels, int dir, const float pos[3], float lod, float out[4]) {
  Pyramid p{levels, R, n_levels, 0};
  texture_lod(p, dir, v3(pos[0], pos[1], pos[2]), lod, out);
}

/* ---------------- voxelize (V1-V5) ---------------- */
/* accum_mode 0: the reference's running average (voxelize.frag:95-120) in canonical fragment order.
 * accum_mode 1: NON-REFERENCE variant (BASELINE.json north_star "deterministic integer or fixed-point atomic accumulation"): every fragment
 *   contributes q_k = (uint)(val_k * 255 + 0.5) per channel, the voxel stores the rounded integer mean (sum_k + n / 2) / n in all 8 bits of
 *   byte k (no count nibble, no 16-sample wrap).  Order independent by construction. */
int orc_voxelize_slab_mode(const orc_scene_t* sc, int R, int z0, int z1, int accum_mode, uint32_t* base, orc_voxel_stats_t* stats) {
  if (!sc || !base || R <= 0) return -1;
  size_t nvox = (size_t)R * R * R;
  std::vector<uint32_t> fx_sum;   /* accum_mode 1: four sums + the count per voxel */
  if (accum_mode >= 1) fx_sum.assign(nvox * 5, 0u);
  memset(base, 0, nvox * (accum_mode == 2 ? sizeof(uint64_t) : sizeof(uint32_t))); /* clear_tex_3d, renderer.cpp:320-321 */
  orc_voxel_stats_t st;
  memset(&st, 0, sizeof st);
  std::vector<uint16_t> per_voxel;
  if (stats) per_voxel.assign(nvox, 0);
  const int W = 2 * R; /* renderer.cpp:339-340 */
  const float fR = (float)R;
  uint32_t nl = sc->n_lights < 10u ? sc->n_lights : 10u;

  for (uint32_t d = 0; d < sc->n_draws; d++) {
    const orc_draw_t& dr = sc->draws[d];
    const orc_material_t& m = sc->mats[dr.material];
    float nm[9];
    normal_matrix(dr.model, nm);
    bool transmissive = (m.illum == 4 || m.illum == 6 || m.illum == 7 || m.illum == 9);
    for (uint32_t t = 0; t + 3 <= dr.index_count; t += 3) {
      V3 wp[3], nn[3];
      for (int k = 0; k < 3; k++) {
        const orc_vertex_t& v = sc->verts[dr.vertex_base + sc->indices[dr.first_index + t + k]];
        V4 w = mat4_mul_point(dr.model, v3(v.pos[0], v.pos[1], v.pos[2]));       /* voxelize.vert:26 */
        wp[k] = v3(w.x / sc->cube_size, w.y / sc->cube_size, w.z / sc->cube_size);
        nn[k] = normalize(mat3_mul(nm, v3(v.norm[0], v.norm[1], v.norm[2])));     /* voxelize.vert:28 */
      }
      int axis = select_axis(wp[0], wp[1], wp[2]);
      float xw[3], yw[3];
      for (int k = 0; k < 3; k++) {
        float a = axis == 1 ? wp[k].y : wp[k].x;
        float b = axis == 0 ? wp[k].y : wp[k].z;
        xw[k] = viewport(a, W); /* R1 */
        yw[k] = viewport(b, W);
      }
      RasterTri rt = raster_setup(xw, yw, W, W);
      uint64_t emitted = 0;
      if (rt.valid) {
        for (int j = rt.jmin; j <= rt.jmax; j++) {
          for (int i = rt.imin; i <= rt.imax; i++) {
            float b[3];
            if (!raster_sample(rt, i, j, b)) continue;
            /* voxelize.frag main() :122-161 */
            V3 pos = v3(interp(b, wp[0].x, wp[1].x, wp[2].x), interp(b, wp[0].y, wp[1].y, wp[2].y),
                        interp(b, wp[0].z, wp[1].z, wp[2].z));
            V3 nrm = v3(interp(b, nn[0].x, nn[1].x, nn[2].x), interp(b, nn[0].y, nn[1].y, nn[2].y),
                        interp(b, nn[0].z, nn[1].z, nn[2].z));
            V3 color = v3(0, 0, 0);
            for (uint32_t li = 0; li < nl; li++) {
              const orc_light_t& L = sc->lights[li];
              V3 lp = divs(v3(L.position[0], L.position[1], L.position[2]), sc->cube_size);
              V3 dv = sub(lp, pos);
              float dist = length(dv);
              V3 dir = divs(dv, dist);
              float att = 1.0f / ((1.0f + 0.0f * dist) + (1.0f * dist) * dist);
              float cos_surf = fmaxf(dot(normalize(nrm), dir), 0.0f);
              float s = cos_surf * att;
              color = add(color, mul(mul(v3(L.color[0], L.color[1], L.color[2]), s), L.intensity));
            }
            color = add(mulv(v3(m.diffuse[0], m.diffuse[1], m.diffuse[2]), color), v3(m.emission[0], m.emission[1], m.emission[2]));
            V3 trans = v3(1, 1, 1);
            float alpha = 1.0f;
            if (transmissive) { trans = v3(m.transmittance[0], m.transmittance[1], m.transmittance[2]); alpha = m.dissolve; }
            float val[4] = {clamp01(trans.x * color.x) * 255.0f, clamp01(trans.y * color.y) * 255.0f,
                            clamp01(trans.z * color.z) * 255.0f, clamp01(alpha) * 255.0f};
            V3 tp = scale_and_bias(pos);
            int vx = (int)(fR * tp.x), vy = (int)(fR * tp.y), vz = (int)(fR * tp.z); /* trunc, R5 */
            if (vx < 0 || vy < 0 || vz < 0 || vx >= R || vy >= R || vz >= R) { st.fragments_oob++; continue; }
            if (vz < z0 || vz >= z1) continue;
            size_t idx = ((size_t)vz * R + vy) * R + vx;
            if (accum_mode >= 1) {
              for (int k = 0; k < 4; k++) fx_sum[idx * 5 + k] += (uint32_t)(val[k] + 0.5f);
              fx_sum[idx * 5 + 4]++;
            } else {
                base[idx] = rgba8_avg_fold(base[idx], val);
            }
            st.fragments++;
            emitted++;
            if (stats && per_voxel[idx] < 65535) per_voxel[idx]++;
          }
        }
      }
      if (!emitted) st.tris_no_frag++;
    }
  }
  if (accum_mode == 1) {
    for (size_t i = 0; i < nvox; i++) {
      const uint32_t n = fx_sum[i * 5 + 4];
      if (!n) continue;
      uint32_t w = 0;
      for (int k = 0; k < 4; k++) w |= ((fx_sum[i * 5 + k] + n / 2u) / n) << (8 * k);
      base[i] = w;
    }
  } else if (accum_mode == 2) { /* RGBA16F grid: the mean colour in [0,1] rounded to half; `base` is really a uint64 array */
    uint64_t* b64 = reinterpret_cast<uint64_t*>(base);
    for (size_t i = 0; i < nvox; i++) {
      const uint32_t n = fx_sum[i * 5 + 4];
      if (!n) continue;
      float c[4];
      for (int k = 0; k < 4; k++) c[k] = (float)fx_sum[i * 5 + k] / ((float)n * 255.0f);
      b64[i] = pack_half4(c);
    }
  }
  if (stats) {
    for (size_t i = 0; i < nvox; i++) {
      if (accum_mode >= 1 ? per_voxel[i] != 0 : base[i] != 0u) st.occupied++;
      if (per_voxel[i] > st.max_per_voxel) st.max_per_voxel = per_voxel[i];
      if (per_voxel[i] >= 16) st.wrapped_voxels++;
    }
    *stats = st;
  }
  return 0;
}

int orc_voxelize_slab(const orc_scene_t* sc, int R, int z0, int z1, uint32_t* base, orc_voxel_stats_t* stats) {
  return orc_voxelize_slab_mode(sc, R, z0, z1, 0, base, stats);
}

int orc_voxelize(const orc_scene_t* sc, int R, uint32_t* base, orc_voxel_stats_t* stats) {
  return orc_voxelize_slab(sc, R, 0, R, base, stats);
}

/* ---------------- mipmap (M1) ---------------- */
int orc_mipmap_fmt(const uint32_t* base, int R, int n_levels, uint32_t* const* out, int fmt);
int orc_mipmap(const uint32_t* base, int R, int n_levels, uint32_t* const* out) { return orc_mipmap_fmt(base, R, n_levels, out, 0); }

/* fmt 1: every array is uint64 per texel (RGBA16F); the blend runs in fp32 on the exactly converted halves, the result is clamped to
 * [0,1] like the unorm store of the reference and rounded to half (nearest even) */
static int g_mip_balanced_sum = 0; /* TEST SWITCH, see orc_debug_set_mip_balanced_sum */
int orc_mipmap_fmt(const uint32_t* base, int R, int n_levels, uint32_t* const* out, int fmt) {
  if (!base || !out || R <= 0 || n_levels < 1) return -1;
  size_t nvox = (size_t)R * R * R;
  const size_t tb = fmt == 1 ? 8 : 4;
  for (int d = 0; d < 6; d++)
    if (out[d * n_levels] && out[d * n_levels] != base) memcpy(out[d * n_levels], base, nvox * tb);
  /* front/back child pairs per direction, children numbered as mipmap.comp:10-20 */
  static const int off[8][3] = {{1, 1, 1}, {1, 1, 0}, {1, 0, 1}, {1, 0, 0}, {0, 1, 1}, {0, 1, 0}, {0, 0, 1}, {0, 0, 0}};
  static const int pairs[6][4][2] = {
      {{0, 4}, {1, 5}, {2, 6}, {3, 7}}, /* -x  mipmap.comp:59-63 */
      {{4, 0}, {5, 1}, {6, 2}, {7, 3}}, /* +x  :66-70 */
      {{0, 2}, {1, 3}, {5, 7}, {4, 6}}, /* -y  :73-77 */
      {{2, 0}, {3, 1}, {7, 5}, {6, 4}}, /* +y  :80-84 */
      {{0, 1}, {2, 3}, {4, 5}, {6, 7}}, /* -z  :87-91 */
      {{1, 0}, {3, 2}, {5, 4}, {7, 6}}, /* +z  :94-98 */
  };
  for (int l = 0; l + 1 < n_levels; l++) {
    int Ns = R >> l; if (Ns < 1) Ns = 1;
    int Nd = R >> (l + 1);
    if (Nd < 1) break; /* non-existent level: the reference's extra dispatches are no-ops */
    for (int d = 0; d < 6; d++) {
      const uint32_t* src = (l == 0) ? base : out[d * n_levels + l];
      uint32_t* dst = out[d * n_levels + l + 1];
#pragma omp parallel for schedule(static)
      for (int z = 0; z < Nd; z++)
        for (int y = 0; y < Nd; y++)
          for (int x = 0; x < Nd; x++) {
            float c[8][4];
            for (int i = 0; i < 8; i++) {
              int sx = 2 * x + off[i][0], sy = 2 * y + off[i][1], sz = 2 * z + off[i][2];
              if (fmt == 0 && unorm_unpack_mode() == 1) unpack_unorm_reciprocal(src[((size_t)sz * Ns + sy) * Ns + sx], c[i]); /* TEST SWITCH */
              else load_texel(src, ((size_t)sz * Ns + sy) * Ns + sx, fmt, c[i]);
            }
            float acc[4];
            for (int k = 0; k < 4; k++) {
              float s = 0.0f, t[4];
              for (int pi = 0; pi < 4; pi++) {
                const float* f = c[pairs[d][pi][0]];
                const float* b = c[pairs[d][pi][1]];
                float v = f[k] + ((1.0f - f[3]) * b[k]); /* alpha_blend :40-43 */
                t[pi] = v;
                s = pi == 0 ? v : s + v;
              }
              if (g_mip_balanced_sum) s = (t[0] + t[1]) + (t[2] + t[3]); /* TEST SWITCH: Mesa's GLSL compiler rebalances the four-term sum */
              acc[k] = s / 4.0f;
            }
            if (fmt == 1) reinterpret_cast<uint64_t*>(dst)[((size_t)z * Nd + y) * Nd + x] = pack_half4(acc);
            else dst[((size_t)z * Nd + y) * Nd + x] = pack_unorm(acc);
          }
    }
  }
  return 0;
}

/* ---------------- camera pass (C1) ---------------- */
static void feed_camera_pass(CameraPass& pass, const orc_scene_t* sc, const float view[16], const float proj[16]) {
  float pv[16];
  mat4_mul(proj, view, pv); /* projection * view, voxel_cone_tracing.vert:25 */
  uint32_t seq = 0;
  for (uint32_t d = 0; d < sc->n_draws; d++) {
    const orc_draw_t& dr = sc->draws[d];
    float nm[9];
    normal_matrix(dr.model, nm);
    for (uint32_t t = 0; t + 3 <= dr.index_count; t += 3, seq++) {
      FFVertex in[3];
      for (int k = 0; k < 3; k++) {
        const orc_vertex_t& v = sc->verts[dr.vertex_base + sc->indices[dr.first_index + t + k]];
        V4 w = mat4_mul_point(dr.model, v3(v.pos[0], v.pos[1], v.pos[2])); /* :24 */
        in[k].world = v3(w.x, w.y, w.z);
        in[k].nrm = normalize(mat3_mul(nm, v3(v.norm[0], v.norm[1], v.norm[2]))); /* :26 */
        in[k].clip = mat4_mul_v4(pv, w);
      }
      pass.add_triangle(in, dr.material, seq);
    }
  }
}

int orc_gbuffer(const orc_scene_t* sc, const float view[16], const float proj[16], int W, int H, uint32_t* tri_id,
                float* depth, float* world_pos, float* normal, uint32_t* material) {
  if (!sc || !tri_id || !depth) return -1;
  CameraPass pass(W, H);   /* clipping, rasterisation, depth test, interpolation: vct_fixed_function.h */
  feed_camera_pass(pass, sc, view, proj);
  pass.resolve(tri_id, depth, world_pos, normal, material);
  return 0;
}

/* ---------------- shading (C2-C6) ---------------- */
int orc_trace(const orc_scene_t* sc, const float view[16], int W, int H, const uint32_t* tri_id, const float* world_pos,
              const float* normal, const uint32_t* material, const uint32_t* const* levels, int R, int n_levels,
              const orc_trace_params_t* prm, int row0, int row1, int tile_stride, int tile_phase, uint32_t* frame,
              orc_trace_stats_t* stats) {
  return orc_trace_fmt(sc, view, W, H, tri_id, world_pos, normal, material, levels, R, n_levels, 0, prm, row0, row1, tile_stride, tile_phase, frame, stats);
}

int orc_trace_fmt(const orc_scene_t* sc, const float view[16], int W, int H, const uint32_t* tri_id, const float* world_pos,
                  const float* normal, const uint32_t* material, const uint32_t* const* levels, int R, int n_levels, int fmt,
                  const orc_trace_params_t* prm, int row0, int row1, int tile_stride, int tile_phase, uint32_t* frame,
                  orc_trace_stats_t* stats) {
  if (!sc || !tri_id || !world_pos || !normal || !material || !levels || !prm || !frame) return -1;
  ShadeCtx c;
  c.scene = sc;
  c.pyr = Pyramid{levels, R, n_levels, fmt};
  c.prm = prm;
  c.camera_position = v3(view[12], view[13], view[14]); /* glm::column(view, 3), renderer.cpp:279 (sic) */
  if (row0 < 0) row0 = 0;
  if (row1 > H || row1 <= 0) row1 = H;
  if (tile_stride < 1) tile_stride = 1;
  int tiles_x = (W + 31) / 32;
  uint64_t n_shaded = 0, s_d = 0, s_sh = 0, s_sp = 0, s_rf = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : n_shaded, s_d, s_sh, s_sp, s_rf)
  for (int j = row0; j < row1; j++) {
    for (int i = 0; i < W; i++) {
      int tile = (j / 32) * tiles_x + (i / 32);
      if (tile % tile_stride != tile_phase) continue;
      size_t px = (size_t)j * W + i;
      if (tri_id[px] == 0xFFFFFFFFu) { frame[px] = 0xFF404026u; continue; } /* R8: (38,64,64,255) */
      SampleCount cnt;
      float rgba[4];
      const orc_material_t& m = sc->mats[material[px]];
      bool wrote = shade(c, m, v3(world_pos[px * 3], world_pos[px * 3 + 1], world_pos[px * 3 + 2]),
                         v3(normal[px * 3], normal[px * 3 + 1], normal[px * 3 + 2]), rgba, cnt);
      if (!wrote) { frame[px] = 0xFF404026u; continue; }
      if (prm->view_voxel_dir < 7) {
        /* blend SRC_ALPHA / ONE_MINUS_SRC_ALPHA over the clear colour (renderer.cpp:387-388) */
        const float bg[4] = {0.15f, 0.25f, 0.25f, 1.0f};
        float a = rgba[3];
        for (int k = 0; k < 4; k++) rgba[k] = rgba[k] * a + bg[k] * (1.0f - a);
      }
      frame[px] = pack_unorm(rgba);
      n_shaded++;
      s_d += cnt.diffuse; s_sh += cnt.shadow; s_sp += cnt.specular; s_rf += cnt.refraction;
    }
  }
  if (stats) {
    stats->shaded_pixels = n_shaded;
    stats->samples_diffuse = s_d; stats->samples_shadow = s_sh; stats->samples_specular = s_sp; stats->samples_refraction = s_rf;
  }
  return 0;
}

/* TEST-ONLY (tests/test_gl_llvmpipe.py): Renderer::visualize as a forward renderer -- every fragment that passes the depth test when it is
 * drawn is shaded and blended (SRC_ALPHA, ONE_MINUS_SRC_ALPHA, renderer.cpp:387-388) into an RGBA8 colour buffer over what is there; a
 * fragment whose shader returns without writing (outside the cube) contributes what Mesa llvmpipe makes of the unwritten output, zero.
 * orc_gbuffer + orc_trace keep one layer per pixel: the same picture whenever alpha is 1 and every nearest fragment writes. */
int orc_render_forward(const orc_scene_t* sc, const float view[16], const float proj[16], int W, int H, const uint32_t* const* levels, int R,
                       int n_levels, const orc_trace_params_t* prm, uint32_t* frame) {
  if (!sc || !levels || !prm || !frame) return -1;
  ShadeCtx c;
  c.scene = sc;
  c.pyr = Pyramid{levels, R, n_levels, 0};
  c.prm = prm;
  c.camera_position = v3(view[12], view[13], view[14]);
  for (size_t i = 0; i < (size_t)W * H; i++) frame[i] = kClearColour;
  CameraPass pass(W, H);
  feed_camera_pass(pass, sc, view, proj);
  pass.forward([&](size_t px, V3 world, V3 nrm, uint32_t material) {
    SampleCount cnt;
    float src[4], dst[4];
    if (!shade(c, sc->mats[material], world, nrm, src, cnt)) src[0] = src[1] = src[2] = src[3] = 0.0f;
    for (int k = 0; k < 4; k++) src[k] = clamp01(src[k]);   /* a fixed-point colour buffer clamps the fragment colour before blending */
    unpack_unorm(frame[px], dst);
    const float a = src[3];
    for (int k = 0; k < 4; k++) dst[k] = src[k] * a + dst[k] * (1.0f - a);
    frame[px] = pack_unorm(dst);
  });
  return 0;
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_debug_set_lod_filter(int mode) { vct_ff::lod_filter_mode() = mode; }
void orc_debug_set_mip_balanced_sum(int on) { g_mip_balanced_sum = on; }
void orc_debug_set_unorm_unpack(int mode) { vct_ff::unorm_unpack_mode() = mode; }
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

} /* extern "C" */
