/*
 * glsl_env.h -- what GLSL 4.50 means beyond C++ syntax, for the reference's shader text translated by glsl2cpp.py.
 * TEST INFRASTRUCTURE ONLY (oracle/).  Included by harness.cpp once per build mode, inside namespace GLREF_NS.
 *
 * Vector / matrix types, swizzles and the built-in functions come from the REFERENCE'S OWN vendored GLM 0.9.9
 * (/root/reference/thirdparty/glm, the library its CPU side already uses; GLM_FORCE_SWIZZLE).  On top of that this header
 * supplies only
 *   (1) the implicit int -> float conversions GLSL performs and C++ template deduction refuses
 *       (`max(x, 0)`, `clamp(v, 0, 1)`, `vec4 / 4`, `ivec3 * vec3`, a swizzle where a vec3 is expected);
 *   (2) the opaque types of the hot path (sampler3D, image3D, uimage3D) and their functions, implemented by the
 *       fixed-function rules the oracle uses (vct_fixed_function.h: R6 texel conversion, R7 textureLod);
 *   (3) two BUILD MODES for the built-ins whose precision / evaluation order GLSL leaves to the implementation:
 *       GLREF_RULES = 0  GLM's own definitions, untouched (normalize = v * inversesqrt(dot), round = half away from zero,
 *                        mat4 * vec4 = (c0 x + c1 y) + (c2 z + c3 w), float inverse());
 *       GLREF_RULES = 1  the oracle's written rule R5 / R9 (normalize = v / sqrt(dot), round = ties to even, left-to-right
 *                        sums, double-precision cofactor inverse) -- with these the reference's shader text must reproduce
 *                        the oracle BIT FOR BIT, which is what tests/test_glsl_ref.py asserts.
 */
#include <array>
#include <cmath>
#include <vector>

namespace GLREF_NS {

using glm::vec2; using glm::vec3; using glm::vec4; using glm::ivec3; using glm::uvec3; using glm::mat3; using glm::mat4;
using glm::uint;
template <class T, size_t N> using glsl_array = std::array<T, N>;
template <int N, class T, glm::qualifier Q, int E0, int E1, int E2, int E3>
using swz = glm::detail::_swizzle<N, T, Q, E0, E1, E2, E3>;

/* built-ins the shaders call that need no help */
using glm::abs; using glm::cross; using glm::transpose; using glm::log2; using glm::pow; using glm::tan; using glm::sqrt;
using glm::min; using glm::max; using glm::clamp; using glm::dot; using glm::length;

/* ---- (1) implicit conversions ---- */
/* max(x, 0), clamp(emission, 0, 1), clamp(vec4, 0, 1).  GLSL leaves min / max / clamp of a NaN undefined (a vertex normal of length
 * zero makes normalize() produce one): the rules build returns the non-NaN operand, as IEEE minNum / maxNum, NVIDIA's FMNMX, CUDA's
 * fminf / fmaxf and the oracle do; the glm build keeps GLM's `(a < b) ? b : a`, which passes the NaN on. */
#if GLREF_RULES
inline float max(float a, float b) { return fmaxf(a, b); }
inline float min(float a, float b) { return fminf(a, b); }
inline float clamp(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
#else
inline float max(float a, float b) { return glm::max(a, b); }
inline float min(float a, float b) { return glm::min(a, b); }
inline float clamp(float v, float lo, float hi) { return glm::clamp(v, lo, hi); }
#endif
inline int min(int a, int b) { return glm::min(a, b); }
inline vec3 clamp(vec3 const& v, float lo, float hi) { return vec3(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi)); }
inline vec4 clamp(vec4 const& v, float lo, float hi) { return vec4(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi), clamp(v.w, lo, hi)); }
inline vec4 operator/(vec4 const& v, float s) { return vec4(v.x / s, v.y / s, v.z / s, v.w / s); }   /* (a + b + c + d) / 4, rval / 2 */
inline vec4 operator*(vec4 const& v, float s) { return vec4(v.x * s, v.y * s, v.z * s, v.w * s); }   /* rval * n (uint) */
inline vec3 operator*(ivec3 const& a, vec3 const& b) { return vec3(a) * b; }                          /* dim * scale_and_bias(pos) */
template <class T, glm::qualifier Q, int E0, int E1, int E2, int E3>
inline vec3 operator/(swz<3, T, Q, E0, E1, E2, E3> const& a, float s) { vec3 v = a; return vec3(v.x / s, v.y / s, v.z / s); }

/* ---- (3) implementation-defined built-ins ---- */
#if GLREF_RULES
inline float dot(vec3 const& a, vec3 const& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float length(vec3 const& a) { return sqrtf(dot(a, a)); }
inline vec3 normalize(vec3 const& a) { float l = length(a); return vec3(a.x / l, a.y / l, a.z / l); }            /* R5 */
inline float distance(vec3 const& a, vec3 const& b) { return length(a - b); }
inline vec4 round(vec4 const& v) { return vec4(rintf(v.x), rintf(v.y), rintf(v.z), rintf(v.w)); }                  /* R5 */
inline vec3 mix(vec3 const& a, vec3 const& b, float t) { return a * (1.0f - t) + b * t; }                          /* GLSL 8.3 */
inline vec3 reflect(vec3 const& I, vec3 const& N) { return I - N * (2.0f * dot(N, I)); }                           /* GLSL 8.5 */
inline vec3 refract(vec3 const& I, vec3 const& N, float eta) {                                                     /* GLSL 8.5 */
  float d = dot(N, I);
  float k = 1.0f - eta * eta * (1.0f - d * d);
  if (k < 0.0f) return vec3(0.0f);
  return I * eta - N * (eta * d + sqrtf(k));
}
inline vec4 operator*(mat4 const& m, vec4 const& v) {   /* ((c0 x + c1 y) + c2 z) + c3 w */
  return ((m[0] * v.x + m[1] * v.y) + m[2] * v.z) + m[3] * v.w;
}
inline vec3 operator*(mat3 const& m, vec3 const& v) { return (m[0] * v.x + m[1] * v.y) + m[2] * v.z; }
inline mat4 operator*(mat4 const& a, mat4 const& b) {
  mat4 r;
  for (int c = 0; c < 4; c++) r[c] = ((a[0] * b[c][0] + a[1] * b[c][1]) + a[2] * b[c][2]) + a[3] * b[c][3];
  return r;
}
/* R9: the shaders only ever use mat3(transpose(inverse(model))) of an affine model matrix; its 3x3 part is
 * cofactor(upper 3x3) / det evaluated in double precision, rounded to float.  The translation column is filled in for
 * completeness (-A^-1 t) and is not used by the hot path. */
inline mat4 inverse(mat4 const& m) {
  double a00 = m[0][0], a10 = m[0][1], a20 = m[0][2];
  double a01 = m[1][0], a11 = m[1][1], a21 = m[1][2];
  double a02 = m[2][0], a12 = m[2][1], a22 = m[2][2];
  double c00 = a11 * a22 - a12 * a21, c01 = a12 * a20 - a10 * a22, c02 = a10 * a21 - a11 * a20;
  double c10 = a02 * a21 - a01 * a22, c11 = a00 * a22 - a02 * a20, c12 = a01 * a20 - a00 * a21;
  double c20 = a01 * a12 - a02 * a11, c21 = a02 * a10 - a00 * a12, c22 = a00 * a11 - a01 * a10;
  double det = (a00 * c00 + a01 * c01) + a02 * c02;
  /* inverse(r, c) = C_cr / det; glm is column-major: r[c][r] */
  mat4 r(1.0f);
  r[0][0] = (float)(c00 / det); r[1][0] = (float)(c10 / det); r[2][0] = (float)(c20 / det);
  r[0][1] = (float)(c01 / det); r[1][1] = (float)(c11 / det); r[2][1] = (float)(c21 / det);
  r[0][2] = (float)(c02 / det); r[1][2] = (float)(c12 / det); r[2][2] = (float)(c22 / det);
  double t0 = m[3][0], t1 = m[3][1], t2 = m[3][2];
  for (int i = 0; i < 3; i++) r[3][i] = (float)(-((double)r[0][i] * t0 + (double)r[1][i] * t1 + (double)r[2][i] * t2));
  return r;
}
#else
using glm::normalize; using glm::round; using glm::inverse;
inline float distance(vec3 const& a, vec3 const& b) { return glm::distance(a, b); }          /* distance(vec3, swizzle) */
inline vec3 mix(vec3 const& a, vec3 const& b, float t) { return glm::mix(a, b, t); }         /* mix(vec3, swizzle, float) */
inline vec3 reflect(vec3 const& I, vec3 const& N) { return glm::reflect(I, N); }
inline vec3 refract(vec3 const& I, vec3 const& N, float eta) { return glm::refract(I, N, eta); }
#endif

/* ---- (2) opaque types ---- */
/* sampler3D: one of the reference's six mipmapped RGBA8 textures (texture_3d.cpp:3-25: LINEAR_MIPMAP_LINEAR, CLAMP_TO_BORDER) */
struct sampler3D {
  const uint32_t* const* levels = nullptr;   /* [n_levels] */
  int R = 0, n_levels = 0;
};
/* PERFORMANCE-MODEL hook (glref_*_tex_model, tools/tex_lane_model.py): when a recorder is installed every textureLod call
 * appends which levels the fetch blends and whether their 2x2x2 footprints hold only zero texels.  Never set during parity checks. */
struct TexRecord { float lod; bool empty_l0, empty_l1; };
inline std::vector<TexRecord>*& tex_recorder() { static thread_local std::vector<TexRecord>* r = nullptr; return r; }
inline bool footprint_is_zero(const uint32_t* tex, int N, vec3 const& p) {
  const float ux = p.x * (float)N - 0.5f, uy = p.y * (float)N - 0.5f, uz = p.z * (float)N - 0.5f;
  if (!(fabsf(ux) < 1.0e9f) || !(fabsf(uy) < 1.0e9f) || !(fabsf(uz) < 1.0e9f)) return true;
  const int ix = (int)floorf(ux), iy = (int)floorf(uy), iz = (int)floorf(uz);
  for (int dz = 0; dz < 2; dz++)
    for (int dy = 0; dy < 2; dy++)
      for (int dx = 0; dx < 2; dx++) {
        const int x = ix + dx, y = iy + dy, z = iz + dz;
        if (x < 0 || y < 0 || z < 0 || x >= N || y >= N || z >= N) continue;
        if (tex[((size_t)z * N + y) * N + x]) return false;
      }
  return true;
}
inline vec4 textureLod(sampler3D const& s, vec3 const& p, float lod) {   /* R7 */
  if (tex_recorder()) {
    const float l = fminf(fmaxf(lod, 0.0f), (float)(s.n_levels - 1));
    const int l0 = (int)floorf(l), l1 = l0 + 1 < s.n_levels ? l0 + 1 : s.n_levels - 1;
    tex_recorder()->push_back(TexRecord{l, footprint_is_zero(s.levels[l0], s.R >> l0, p), footprint_is_zero(s.levels[l1], s.R >> l1, p)});
  }
  vct_ff::Pyramid pyr{s.levels, s.R, s.n_levels, 0};
  float out[4];
  vct_ff::texture_lod(pyr, 0, vct_ff::v3(p.x, p.y, p.z), lod, out);
  return vec4(out[0], out[1], out[2], out[3]);
}
inline vec4 texelFetch(sampler3D const& s, ivec3 const& p, int lod) {    /* R6: unorm8 -> float */
  int N = s.R >> lod;
  float c[4];
  vct_ff::unpack_unorm(s.levels[lod][((size_t)p.z * N + p.y) * N + p.x], c);
  return vec4(c[0], c[1], c[2], c[3]);
}
/* image3D: one RGBA8 level bound for writing (renderer.cpp:297-301) */
struct image3D {
  uint32_t* texels = nullptr;
  int N = 0;
};
inline void imageStore(image3D const& img, ivec3 const& p, vec4 const& v) {   /* R6: float -> unorm8 */
  float c[4] = {v.x, v.y, v.z, v.w};
  img.texels[((size_t)p.z * img.N + p.y) * img.N + p.x] = vct_ff::pack_unorm(c);
}
/* uimage3D: level 0 aliased as r32ui (renderer.cpp:331) */
struct uimage3D {
  uint32_t* texels = nullptr;
  int N = 0;
};
inline ivec3 imageSize(uimage3D const& img) { return ivec3(img.N); }
inline uint imageAtomicCompSwap(uimage3D const& img, ivec3 const& p, uint compare, uint data) {
  /* out-of-bounds image accesses are discarded / return 0 (OpenGL 4.5 8.26) */
  if (p.x < 0 || p.y < 0 || p.z < 0 || p.x >= img.N || p.y >= img.N || p.z >= img.N) return 0u;
  uint32_t& t = img.texels[((size_t)p.z * img.N + p.y) * img.N + p.x];
  uint old = t;
  if (old == compare) t = data;
  return old;
}

}  // namespace GLREF_NS
