/*
 * vct_fixed_function.h -- the FIXED-FUNCTION half of the CPU oracle: everything OpenGL does between and around the shaders
 * (viewport transform, rasterisation, attribute interpolation, near-plane clipping, depth test, texel conversion, textureLod),
 * as the written rules R1-R4, R6-R8 of vct_oracle.cpp.  TEST INFRASTRUCTURE ONLY.
 *
 * Two programs include it: vct_oracle.cpp (the restated shaders) and glsl_ref/harness.cpp (the reference's own GLSL text,
 * translated mechanically and compiled against the reference's GLM).  Both therefore differ ONLY in the programmable stages.
 */
#ifndef VCT_FIXED_FUNCTION_H
#define VCT_FIXED_FUNCTION_H

#include <math.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <vector>

namespace vct_ff {

struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 mul(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 mulv(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 divs(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
inline V3 neg(V3 a) { return v3(-a.x, -a.y, -a.z); }
inline float clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

/* R1: viewport transform of one normalised device coordinate onto an axis of `extent` pixels */
inline float viewport(float ndc, int extent) { return (ndc + 1.0f) * ((float)extent * 0.5f); }

/* ------------------------------------------------------------------ */
/* R2/R3 rasteriser */
struct RasterTri {
  int64_t X[3], Y[3];
  int64_t area;  /* > 0 after orientation fix */
  int sign;
  int imin, imax, jmin, jmax;
  bool valid;
};

inline int64_t floor_div(int64_t a, int64_t b) { int64_t q = a / b, r = a % b; return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q; }
inline int64_t ceil_div(int64_t a, int64_t b) { return -floor_div(-a, b); }

inline RasterTri raster_setup(const float xw[3], const float yw[3], int W, int H) {
  RasterTri t;
  t.valid = false;
  for (int k = 0; k < 3; k++) {
    if (!(fabsf(xw[k]) <= 2097152.0f) || !(fabsf(yw[k]) <= 2097152.0f)) return t; /* guard band, also NaN */
    t.X[k] = (int64_t)rintf(xw[k] * 256.0f);
    t.Y[k] = (int64_t)rintf(yw[k] * 256.0f);
  }
  int64_t a = (t.X[1] - t.X[0]) * (t.Y[2] - t.Y[0]) - (t.Y[1] - t.Y[0]) * (t.X[2] - t.X[0]);
  if (a == 0) return t;
  t.sign = a > 0 ? 1 : -1;
  t.area = a > 0 ? a : -a;
  int64_t minx = std::min(t.X[0], std::min(t.X[1], t.X[2])), maxx = std::max(t.X[0], std::max(t.X[1], t.X[2]));
  int64_t miny = std::min(t.Y[0], std::min(t.Y[1], t.Y[2])), maxy = std::max(t.Y[0], std::max(t.Y[1], t.Y[2]));
  int64_t i0 = ceil_div(minx - 128, 256), i1 = floor_div(maxx - 128, 256);
  int64_t j0 = ceil_div(miny - 128, 256), j1 = floor_div(maxy - 128, 256);
  if (i0 < 0) i0 = 0;
  if (j0 < 0) j0 = 0;
  if (i1 > W - 1) i1 = W - 1;
  if (j1 > H - 1) j1 = H - 1;
  if (i0 > i1 || j0 > j1) return t;
  t.imin = (int)i0; t.imax = (int)i1; t.jmin = (int)j0; t.jmax = (int)j1;
  t.valid = true;
  return t;
}

/* coverage + barycentrics of pixel (i,j); returns false if not covered */
inline bool raster_sample(const RasterTri& t, int i, int j, float b[3]) {
  int64_t px = (int64_t)i * 256 + 128, py = (int64_t)j * 256 + 128;
  int64_t E[3];
  for (int k = 0; k < 3; k++) {
    int a = (k + 1) % 3, c = (k + 2) % 3;
    int64_t dx = t.X[c] - t.X[a], dy = t.Y[c] - t.Y[a];
    int64_t e = dx * (py - t.Y[a]) - dy * (px - t.X[a]);
    if (t.sign < 0) { e = -e; dx = -dx; dy = -dy; }
    if (e < 0) return false;
    if (e == 0) {
      bool topleft = (dy < 0) || (dy == 0 && dx < 0);
      if (!topleft) return false;
    }
    E[k] = e;
  }
  float fa = (float)t.area;
  b[0] = (float)E[0] / fa;
  b[1] = (float)E[1] / fa;
  b[2] = (float)E[2] / fa;
  return true;
}

inline float interp(const float b[3], float a0, float a1, float a2) { return (b[0] * a0 + b[1] * a1) + b[2] * a2; }

/* ------------------------------------------------------------------ */
/* textures (R6, R7) */
/* Texel storage: fmt 0 = RGBA8 unorm (one uint32 per texel, the reference's format, texture_3d.cpp:3-25);
 * fmt 1 = RGBA16F (one uint64 per texel: four IEEE halves, R in the low 16 bits) -- the NON-REFERENCE storage variant of BASELINE.json
 * config 5 ("fp16 RGBA + full mip chain").  The level pointers are then really uint64 arrays. */
struct Pyramid {
  const uint32_t* const* levels; /* [dir * n_levels + level] */
  int R, n_levels;
  int fmt = 0;
  inline int size(int l) const { int n = R >> l; return n < 1 ? 1 : n; }
};
inline float half_to_float(uint16_t h) { _Float16 v; memcpy(&v, &h, 2); return (float)v; }
inline uint16_t float_to_half(float f) { _Float16 v = (_Float16)f; uint16_t h; memcpy(&h, &v, 2); return h; } /* round to nearest even */
inline void unpack_half4(uint64_t c, float out[4]) {
  for (int k = 0; k < 4; k++) out[k] = half_to_float((uint16_t)(c >> (16 * k)));
}
inline uint64_t pack_half4(const float v[4]) {
  uint64_t r = 0;
  for (int k = 0; k < 4; k++) r |= (uint64_t)float_to_half(clamp01(v[k])) << (16 * k);
  return r;
}

/* TEST SWITCH (tests/test_gl_llvmpipe.py only; 0 everywhere else; read by the oracle's mip filter alone): 1 = unorm8 -> float as Mesa llvmpipe
 * converts a texel, c * (1.0f / 255.0f) -- one rounding more than the specification's c / 255 (rule R6), up to one ulp off, which moves results
 * of the mip filter that sit exactly on a rounding tie (a quarter of a sum of 8-bit values often does). */
inline int g_unorm_unpack_mode = 0;   /* C++17 inline variable: one per shared library */
inline int& unorm_unpack_mode() { return g_unorm_unpack_mode; }
inline void unpack_unorm_reciprocal(uint32_t c, float out[4]) {
  const float r = 1.0f / 255.0f;
  for (int k = 0; k < 4; k++) out[k] = (float)((c >> (8 * k)) & 0xFFu) * r;
}
inline void unpack_unorm(uint32_t c, float out[4]) {
  out[0] = (float)(c & 0xFFu) / 255.0f;
  out[1] = (float)((c >> 8) & 0xFFu) / 255.0f;
  out[2] = (float)((c >> 16) & 0xFFu) / 255.0f;
  out[3] = (float)((c >> 24) & 0xFFu) / 255.0f;
}
inline uint32_t pack_unorm(const float v[4]) {
  uint32_t r = 0;
  for (int k = 0; k < 4; k++) r |= ((uint32_t)rintf(clamp01(v[k]) * 255.0f)) << (8 * k);
  return r;
}

inline void load_texel(const uint32_t* tex, size_t idx, int fmt, float c[4]);
inline void trilinear(const uint32_t* tex, int N, V3 s, float out[4], int fmt = 0) {
  out[0] = out[1] = out[2] = out[3] = 0.0f;
  float ux = s.x * (float)N - 0.5f, uy = s.y * (float)N - 0.5f, uz = s.z * (float)N - 0.5f;
  if (!(fabsf(ux) < 1.0e9f) || !(fabsf(uy) < 1.0e9f) || !(fabsf(uz) < 1.0e9f)) return; /* NaN / absurd: border */
  float fx = floorf(ux), fy = floorf(uy), fz = floorf(uz);
  int ix = (int)fx, iy = (int)fy, iz = (int)fz;
  float ax = ux - fx, ay = uy - fy, az = uz - fz;
  for (int dz = 0; dz < 2; dz++) {
    int z = iz + dz;
    if (z < 0 || z >= N) continue;
    float wz = dz ? az : 1.0f - az;
    for (int dy = 0; dy < 2; dy++) {
      int y = iy + dy;
      if (y < 0 || y >= N) continue;
      float wy = dy ? ay : 1.0f - ay;
      for (int dx = 0; dx < 2; dx++) {
        int x = ix + dx;
        if (x < 0 || x >= N) continue;
        float wx = dx ? ax : 1.0f - ax;
        float w = (wx * wy) * wz;
        float c[4];
        load_texel(tex, ((size_t)z * N + y) * N + x, fmt, c);
        for (int k = 0; k < 4; k++) out[k] = out[k] + w * c[k];
      }
    }
  }
}

inline void load_texel(const uint32_t* tex, size_t idx, int fmt, float c[4]) {
  if (fmt == 1) unpack_half4(reinterpret_cast<const uint64_t*>(tex)[idx], c);
  else unpack_unorm(tex[idx], c);
}

/* TEST SWITCH (tests/test_gl_llvmpipe.py only; 0 everywhere else): 1 = the mip filter of Mesa llvmpipe, the one OpenGL implementation this
 * image has.  llvmpipe does not blend two levels linearly in the LOD fraction as the GL specification writes it (rule R7) but "brilinearly"
 * (gallivm lp_bld_sample.c, lp_build_brilinear_lod with BRILINEAR_FACTOR 2): lod + 0.25 is split into level and fraction phi, the weight of
 * the coarser level is 2 phi - 1, and a weight <= 0 means the finer level alone -- i.e. one level for fractions below 0.25 and above 0.75, a
 * ramp of twice the slope between.  With the switch on, the oracle can be compared with the reference's shaders RUNNING on llvmpipe without
 * the driver's shortcut drowning everything else. */
inline int g_lod_filter_mode = 0;
inline int& lod_filter_mode() { return g_lod_filter_mode; }

inline void texture_lod(const Pyramid& p, int dir, V3 s, float lod, float out[4]) {
  float maxl = (float)(p.n_levels - 1);
  float l = fminf(fmaxf(lod, 0.0f), maxl);
  if (!(l == l)) l = 0.0f;
  int l0 = (int)floorf(l);
  int l1 = l0 + 1 < p.n_levels ? l0 + 1 : p.n_levels - 1;
  float f = l - (float)l0;
  if (lod_filter_mode() == 1) {
    float lb = l + 0.25f;
    l0 = (int)floorf(lb);
    f = (lb - (float)l0) * 2.0f + -1.0f;
    if (l0 >= p.n_levels - 1) { l0 = p.n_levels - 1; f = 0.0f; }
    if (f <= 0.0f) f = 0.0f;
    l1 = l0 + 1 < p.n_levels ? l0 + 1 : p.n_levels - 1;
  }
  float t0[4], t1[4];
  trilinear(p.levels[dir * p.n_levels + l0], p.size(l0), s, t0, p.fmt);
  trilinear(p.levels[dir * p.n_levels + l1], p.size(l1), s, t1, p.fmt);
  for (int k = 0; k < 4; k++) out[k] = (1.0f - f) * t0[k] + f * t1[k];
}


/* ------------------------------------------------------------------ */
/* camera pass (C1), fixed-function part: near-plane clipping (R2c), perspective divide + viewport (R1), rasterisation (R2),
 * GL_LESS depth test (R8) and perspective-correct interpolation of the vertex-shader outputs (R3).  The caller runs the vertex
 * stage (voxel_cone_tracing.vert) and hands every triangle over in draw order. */
struct FFVertex { V4 clip; V3 world, nrm; };   /* gl_Position, vs_out.world_position.xyz, vs_out.normal */

class CameraPass {
 public:
  CameraPass(int W_, int H_) : W(W_), H(H_) {}

  /* one triangle of the index buffer; seq = its global sequence number (draw order, then index-buffer order) */
  void add_triangle(const FFVertex in[3], uint32_t material, uint32_t seq) {
    float dn[3]; /* signed distance to the near plane in clip space: z_c + w_c >= 0 is inside (-w <= z) */
    int n_out = 0;
    for (int k = 0; k < 3; k++) {
      dn[k] = in[k].clip.z + in[k].clip.w;
      if (!(dn[k] >= 0.0f)) n_out++;   /* also NaN */
    }
    if (n_out == 3) return;
    /* R2c: near-plane clipping (GL clips primitives against -w <= z, OpenGL 4.5 13.7; src/renderer.cpp:384-388 enables nothing that
     * would change it).  Sutherland-Hodgman on the one plane; a new vertex is always computed FROM THE INSIDE vertex of its edge
     * (t = d_in / (d_in - d_out), P = in + t * (out - in), fp32, no FMA) so that two triangles sharing the edge get the same point.
     * The 3- or 4-vertex polygon is drawn as the fan (p0,p1,p2), (p0,p2,p3); every piece keeps the triangle's sequence number. */
    FFVertex poly[4];
    int np = 0;
    if (n_out == 0) {
      poly[0] = in[0]; poly[1] = in[1]; poly[2] = in[2]; np = 3;
    } else {
      for (int k = 0; k < 3; k++) {
        const int k1 = (k + 1) % 3;
        const bool in_a = dn[k] >= 0.0f, in_b = dn[k1] >= 0.0f;
        if (in_a) poly[np++] = in[k];
        if (in_a != in_b) {
          const FFVertex& vi = in_a ? in[k] : in[k1];
          const FFVertex& vo = in_a ? in[k1] : in[k];
          const float di = in_a ? dn[k] : dn[k1], dout = in_a ? dn[k1] : dn[k];
          const float tt = di / (di - dout);
          FFVertex c;
          c.clip.x = vi.clip.x + tt * (vo.clip.x - vi.clip.x); c.clip.y = vi.clip.y + tt * (vo.clip.y - vi.clip.y);
          c.clip.z = vi.clip.z + tt * (vo.clip.z - vi.clip.z); c.clip.w = vi.clip.w + tt * (vo.clip.w - vi.clip.w);
          c.world = add(vi.world, mul(sub(vo.world, vi.world), tt));
          c.nrm = add(vi.nrm, mul(sub(vo.nrm, vi.nrm), tt));
          poly[np++] = c;
        }
      }
    }
    for (int piece = 0; piece + 3 <= np; piece++) {
      const FFVertex* pvt[3] = {&poly[0], &poly[piece + 1], &poly[piece + 2]};
      TriRec r;
      r.seq = seq;
      r.material = material;
      float xw[3], yw[3];
      bool ok = true;
      for (int k = 0; k < 3; k++) {
        const V4 clip = pvt[k]->clip;
        r.world[k] = pvt[k]->world;
        r.nrm[k] = pvt[k]->nrm;
        if (!(clip.w > 0.0f)) { ok = false; break; } /* R2: cannot happen behind a near plane with near > 0; guards general matrices */
        float iw = 1.0f / clip.w;
        r.iw[k] = iw;
        float xn = clip.x * iw, yn = clip.y * iw, zn = clip.z * iw;
        xw[k] = viewport(xn, W);
        yw[k] = viewport(yn, H);
        r.zw[k] = (zn + 1.0f) * 0.5f;
      }
      if (!ok) continue;
      r.rt = raster_setup(xw, yw, W, H);
      if (!r.rt.valid) continue;
      tris.push_back(r);
    }
  }

  /* depth-tested visibility + interpolated attributes; row 0 = bottom.  Any of world_pos / normal / material may be null. */
  void resolve(uint32_t* tri_id, float* depth, float* world_pos, float* normal, uint32_t* material) const {
    const size_t npx = (size_t)W * H;
    for (size_t i = 0; i < npx; i++) { tri_id[i] = 0xFFFFFFFFu; depth[i] = 1.0f; } /* glClear depth = 1 */
    /* row bands in parallel; inside a band triangles are visited in draw order => GL_LESS, first wins ties */
#pragma omp parallel for schedule(dynamic, 8)
    for (int j = 0; j < H; j++) {
      for (size_t ti = 0; ti < tris.size(); ti++) {
        const TriRec& r = tris[ti];
        if (j < r.rt.jmin || j > r.rt.jmax) continue;
        for (int i = r.rt.imin; i <= r.rt.imax; i++) {
          float b[3];
          if (!raster_sample(r.rt, i, j, b)) continue;
          float zw = interp(b, r.zw[0], r.zw[1], r.zw[2]);
          if (!(zw >= 0.0f && zw <= 1.0f)) continue;
          size_t px = (size_t)j * W + i;
          if (!(zw < depth[px])) continue;
          depth[px] = zw;
          tri_id[px] = r.seq;
          float q[3] = {b[0] * r.iw[0], b[1] * r.iw[1], b[2] * r.iw[2]};
          float qs = (q[0] + q[1]) + q[2];
          if (world_pos) {
            world_pos[px * 3 + 0] = interp(q, r.world[0].x, r.world[1].x, r.world[2].x) / qs;
            world_pos[px * 3 + 1] = interp(q, r.world[0].y, r.world[1].y, r.world[2].y) / qs;
            world_pos[px * 3 + 2] = interp(q, r.world[0].z, r.world[1].z, r.world[2].z) / qs;
          }
          if (normal) {
            normal[px * 3 + 0] = interp(q, r.nrm[0].x, r.nrm[1].x, r.nrm[2].x) / qs;
            normal[px * 3 + 1] = interp(q, r.nrm[0].y, r.nrm[1].y, r.nrm[2].y) / qs;
            normal[px * 3 + 2] = interp(q, r.nrm[0].z, r.nrm[1].z, r.nrm[2].z) / qs;
          }
          if (material) material[px] = r.material;
        }
      }
    }
  }

  /* TEST-ONLY companion of resolve() (tests/test_gl_llvmpipe.py): FORWARD rendering as a GL pipeline does it -- every fragment that passes
   * the depth test at the moment it is drawn is handed to `on_fragment(px, world, normal, material)`, per pixel in draw order, so that the
   * caller can shade it and blend it over what earlier fragments left (the voxel debug view's alpha blending, fragments whose shader
   * writes nothing).  resolve() keeps only the nearest fragment; for frames of alpha 1 the two are the same picture. */
  template <class F>
  void forward(F&& on_fragment) const {
    std::vector<float> depth((size_t)W * H, 1.0f);
#pragma omp parallel for schedule(dynamic, 8)
    for (int j = 0; j < H; j++) {
      for (size_t ti = 0; ti < tris.size(); ti++) {
        const TriRec& r = tris[ti];
        if (j < r.rt.jmin || j > r.rt.jmax) continue;
        for (int i = r.rt.imin; i <= r.rt.imax; i++) {
          float b[3];
          if (!raster_sample(r.rt, i, j, b)) continue;
          float zw = interp(b, r.zw[0], r.zw[1], r.zw[2]);
          if (!(zw >= 0.0f && zw <= 1.0f)) continue;
          size_t px = (size_t)j * W + i;
          if (!(zw < depth[px])) continue;
          depth[px] = zw;
          float q[3] = {b[0] * r.iw[0], b[1] * r.iw[1], b[2] * r.iw[2]};
          float qs = (q[0] + q[1]) + q[2];
          V3 wp, nn;
          wp.x = interp(q, r.world[0].x, r.world[1].x, r.world[2].x) / qs;
          wp.y = interp(q, r.world[0].y, r.world[1].y, r.world[2].y) / qs;
          wp.z = interp(q, r.world[0].z, r.world[1].z, r.world[2].z) / qs;
          nn.x = interp(q, r.nrm[0].x, r.nrm[1].x, r.nrm[2].x) / qs;
          nn.y = interp(q, r.nrm[0].y, r.nrm[1].y, r.nrm[2].y) / qs;
          nn.z = interp(q, r.nrm[0].z, r.nrm[1].z, r.nrm[2].z) / qs;
          on_fragment(px, wp, nn, r.material);
        }
      }
    }
  }

 private:
  struct TriRec { RasterTri rt; V3 world[3], nrm[3]; float iw[3], zw[3]; uint32_t material; uint32_t seq; };
  int W, H;
  std::vector<TriRec> tris;
};

/* R8: the cleared colour buffer (0.15, 0.25, 0.25, 1) as RGBA8 = (38, 64, 64, 255) */
const uint32_t kClearColour = 0xFF404026u;

}  // namespace vct_ff
#endif
