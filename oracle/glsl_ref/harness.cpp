/*
 * harness.cpp -- runs THE REFERENCE'S OWN GLSL (shader/voxelize.{vert,geom,frag}, mipmap.comp, voxel_cone_tracing.{vert,frag})
 * on the CPU.  TEST INFRASTRUCTURE ONLY: the second, reference-sourced checker beside the restated oracle (vct_oracle.cpp).
 *
 * How: oracle/glsl_ref/glsl2cpp.py rewrites the shader text, read from /root/reference where it lies, into C++ member
 * declarations (purely syntactic, see its header; output under oracle/_ref/gen/ during the build only, never committed); each shader becomes the
 * body of a struct below; vectors, matrices, swizzles and built-ins are the reference's vendored GLM 0.9.9
 * (thirdparty/glm) plus glsl_env.h.  This file is the "driver + fixed-function GPU" around them:
 *   * the uniform / vertex-array set-up and draw loop of src/renderer.cpp:241-256 (draw_models), :283-314 (filter),
 *     :316-353 (voxelize), :355-390 (visualize), :258-281 (lights, camera);
 *   * rasterisation, interpolation, clipping, depth test, texel conversion and textureLod from vct_fixed_function.h --
 *     the SAME written rules R1-R4, R6-R8 the oracle uses, so the two programs differ only in the programmable stages.
 * Fragments are executed sequentially in the canonical order R4, so the imageAtomicCompSwap loop of voxelize.frag:95-120
 * runs exactly as written (first swap fails against a non-empty voxel, one averaging step, second swap succeeds).
 *
 * Compiled twice into one library (oracle/Makefile): GLREF_RULES=1 -> glref_rules_* entry points (built-ins follow the
 * oracle's rules R5 / R9: results must equal the oracle's BIT FOR BIT), GLREF_RULES=0 -> glref_glm_* (GLM's own built-ins:
 * shows how far implementation-defined precision moves the result).
 */
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <limits>
#include <vector>

#define GLM_FORCE_SWIZZLE
#include <glm/glm.hpp>
#include <glm/gtc/matrix_access.hpp>

#include "../vct_fixed_function.h"
#include "../vct_oracle.h"   /* POD scene types only */

#if GLREF_RULES
#define GLREF_NS glref_rules
#define GLREF_FN(name) glref_rules_##name
#else
#define GLREF_NS glref_glm
#define GLREF_FN(name) glref_glm_##name
#endif

#include "glsl_env.h"

namespace GLREF_NS {

/* ---- the six shaders: generated member declarations wrapped in one struct each ---- */
struct voxelize_vert {
  vec4 gl_Position;
#include "gen/voxelize_vert.inc"
};
struct voxelize_geom {
  vec4 gl_Position;
  virtual void EmitVertex() = 0;
  void EndPrimitive() {}
  virtual ~voxelize_geom() {}
#include "gen/voxelize_geom.inc"
};
struct voxelize_frag {
#include "gen/voxelize_frag.inc"
};
struct mipmap_comp {
  uvec3 gl_GlobalInvocationID;
#include "gen/mipmap_comp.inc"
};
struct cone_vert {
  vec4 gl_Position;
#include "gen/voxel_cone_tracing_vert.inc"
};
struct cone_frag {
#include "gen/voxel_cone_tracing_frag.inc"
};

struct GeomRun : voxelize_geom {
  struct Out { vec4 pos; GS_OUT v; };
  std::vector<Out> emitted;
  void EmitVertex() override { emitted.push_back(Out{gl_Position, gs_out}); }
};

inline mat4 load_mat4(const float* m) {
  mat4 r;
  for (int c = 0; c < 4; c++)
    for (int k = 0; k < 4; k++) r[c][k] = m[c * 4 + k];
  return r;
}
inline vec3 load3(const float* p) { return vec3(p[0], p[1], p[2]); }

/* material UBO (src/renderer.h:94-120) */
template <class S>
void bind_material(S& s, const orc_material_t& m) {
  s.ambient = load3(m.ambient); s.diffuse = load3(m.diffuse); s.specular = load3(m.specular);
  s.transmittance = load3(m.transmittance); s.emission = load3(m.emission);
  s.shininess = m.shininess; s.ior = m.ior; s.dissolve = m.dissolve; s.illum = m.illum;
  s.roughness = m.roughness; s.metallic = m.metallic; s.sheen = m.sheen; s.clearcoat_thickness = m.clearcoat_thickness;
  s.clearcoat_roughness = m.clearcoat_roughness; s.anisotropy = m.anisotropy; s.anisotropy_rotation = m.anisotropy_rotation;
}
/* upload_lights, src/renderer.cpp:258-273: every queued light is uploaded, the shader clamps the count */
template <class S>
void upload_lights(S& s, const orc_scene_t* sc) {
  for (uint32_t i = 0; i < sc->n_lights && i < s.point_lights.size(); i++) {
    s.point_lights[i].position = load3(sc->lights[i].position);
    s.point_lights[i].color = load3(sc->lights[i].color);
    s.point_lights[i].intensity = sc->lights[i].intensity;
  }
  s.point_light_count = (int)sc->n_lights;
}

/* ---- Renderer::voxelize(), src/renderer.cpp:316-353, without the filter() call ----
 * order / bary: SENSITIVITY STUDY switches (tools/fixed_function_sensitivity.py), both 0 for every parity check:
 *   order 0 = rule R4 (draw order, index-buffer order, pixel row ascending, column ascending) -- fragments executed as they are generated;
 *         1 = rows and columns descending inside every triangle; 2 = triangles in reverse order; 3 = all fragments of the frame in a
 *         seeded random order (GL guarantees no order between the fragments of a draw call: each is one valid execution);
 *   bary  0 = rule R3 (barycentrics from the snapped vertex positions); 1 = from the unsnapped float positions (SURVEY appendix A's
 *         first draft), coverage unchanged. */
struct VoxTriJob { GeomRun::Out e[3]; vct_ff::RasterTri rt; float xw[3], yw[3]; uint32_t material; };
struct VoxFragJob { uint32_t tri; int i, j; };

inline void run_voxel_fragment(voxelize_frag& fs, const VoxTriJob& tj, int i, int j, int bary) {
  using namespace vct_ff;
  float b[3];
  if (!raster_sample(tj.rt, i, j, b)) return;
  if (bary == 1) {
    const float px = (float)i + 0.5f, py = (float)j + 0.5f;
    float e[3];
    for (int k = 0; k < 3; k++) {
      const int a = (k + 1) % 3, c = (k + 2) % 3;
      e[k] = (tj.xw[c] - tj.xw[a]) * (py - tj.yw[a]) - (tj.yw[c] - tj.yw[a]) * (px - tj.xw[a]);
    }
    const float sum = (e[0] + e[1]) + e[2];
    for (int k = 0; k < 3; k++) b[k] = e[k] / sum;
  }
  const GeomRun::Out* e = tj.e;
  for (int c = 0; c < 4; c++) fs.gs_out.world_position[c] = interp(b, e[0].v.world_position[c], e[1].v.world_position[c], e[2].v.world_position[c]);
  for (int c = 0; c < 3; c++) fs.gs_out.normal[c] = interp(b, e[0].v.normal[c], e[1].v.normal[c], e[2].v.normal[c]);
  for (int c = 0; c < 2; c++) fs.gs_out.uv[c] = interp(b, e[0].v.uv[c], e[1].v.uv[c], e[2].v.uv[c]);
  fs._init_globals();
  fs.main();
}

int run_voxelize(const orc_scene_t* sc, int R, uint32_t* const tex[6], uint64_t* n_fragments, int order, int bary, uint32_t seed) {
  using namespace vct_ff;
  const size_t nvox = (size_t)R * R * R;
  for (int i = 0; i < 6; i++) memset(tex[i], 0, nvox * sizeof(uint32_t));   /* clear_tex_3d :320-321 */
  voxelize_vert vs;
  GeomRun gs;
  voxelize_frag fs;
  vs.cube_size = sc->cube_size;      /* :326 */
  fs.cube_size = sc->cube_size;
  /* the camera UBO is bound (:333) but gl_Position of the vertex stage is replaced by the geometry stage */
  vs.projection = mat4(1.0f); vs.view = mat4(1.0f);
  for (int i = 0; i < 6; i++) { fs.tex3D[i].texels = tex[i]; fs.tex3D[i].N = R; }   /* :330-331 */
  upload_lights(fs, sc);
  const int W = 2 * R;               /* :339-340 */
  uint64_t frags = 0;
  std::vector<VoxTriJob> tri_jobs;   /* order != 0 only */
  std::vector<VoxFragJob> frag_jobs;
  for (uint32_t d = 0; d < sc->n_draws; d++) {       /* draw_models :241-256 */
    const orc_draw_t& dr = sc->draws[d];
    vs.model = load_mat4(dr.model);
    bind_material(fs, sc->mats[dr.material]);
    for (uint32_t t = 0; t + 3 <= dr.index_count; t += 3) {
      for (int k = 0; k < 3; k++) {
        const orc_vertex_t& v = sc->verts[dr.vertex_base + sc->indices[dr.first_index + t + k]];
        vs.position = load3(v.pos); vs.normal = load3(v.norm); vs.uv = vec2(v.uv[0], v.uv[1]);
        vs._init_globals();
        vs.main();
        gs.vs_out[k].world_position = vs.vs_out.world_position;
        gs.vs_out[k].normal = vs.vs_out.normal;
        gs.vs_out[k].uv = vs.vs_out.uv;
      }
      gs.emitted.clear();
      gs._init_globals();
      gs.main();
      if (gs.emitted.size() != 3) return -2;
      /* fixed function: divide by w (= 1), viewport R1, rasterise R2, interpolate R3 (w = 1: affine) */
      VoxTriJob tj;
      for (int k = 0; k < 3; k++) {
        tj.e[k] = gs.emitted[k];
        const vec4 p = gs.emitted[k].pos;
        tj.xw[k] = viewport(p.x / p.w, W);
        tj.yw[k] = viewport(p.y / p.w, W);
      }
      tj.material = dr.material;
      tj.rt = raster_setup(tj.xw, tj.yw, W, W);
      if (!tj.rt.valid) continue;
      for (int j = tj.rt.jmin; j <= tj.rt.jmax; j++)
        for (int i = tj.rt.imin; i <= tj.rt.imax; i++) {
          float b[3];
          if (!raster_sample(tj.rt, i, j, b)) continue;
          frags++;
          if (order == 0) run_voxel_fragment(fs, tj, i, j, bary);
          else frag_jobs.push_back(VoxFragJob{(uint32_t)tri_jobs.size(), i, j});
        }
      if (order != 0) tri_jobs.push_back(tj);
    }
  }
  if (order != 0) {
    const size_t n = frag_jobs.size();
    if (order == 1) {          /* inside every triangle: last fragment first */
      size_t a = 0;
      while (a < n) {
        size_t b = a;
        while (b < n && frag_jobs[b].tri == frag_jobs[a].tri) b++;
        std::reverse(frag_jobs.begin() + a, frag_jobs.begin() + b);
        a = b;
      }
    } else if (order == 2) {   /* last triangle first, fragments of a triangle in rule order */
      std::vector<VoxFragJob> r;
      r.reserve(n);
      size_t b = n;
      while (b > 0) {
        size_t a = b - 1;
        while (a > 0 && frag_jobs[a - 1].tri == frag_jobs[b - 1].tri) a--;
        r.insert(r.end(), frag_jobs.begin() + a, frag_jobs.begin() + b);
        b = a;
      }
      frag_jobs.swap(r);
    } else {                    /* Fisher-Yates with a 64-bit LCG */
      uint64_t st = 0x9E3779B97F4A7C15ull ^ seed;
      for (size_t k = n; k > 1; k--) {
        st = st * 6364136223846793005ull + 1442695040888963407ull;
        const size_t r = (size_t)((st >> 33) % k);
        std::swap(frag_jobs[k - 1], frag_jobs[r]);
      }
    }
    uint32_t bound = 0xFFFFFFFFu;
    for (const VoxFragJob& fj : frag_jobs) {
      const VoxTriJob& tj = tri_jobs[fj.tri];
      if (tj.material != bound) { bind_material(fs, sc->mats[tj.material]); bound = tj.material; }
      run_voxel_fragment(fs, tj, fj.i, fj.j, bary);
    }
  }
  if (n_fragments) *n_fragments = frags;
  return 0;
}

/* ---- Renderer::filter(), src/renderer.cpp:283-314 ---- */
int run_mipmap(uint32_t* const* levels, int R, int n_levels) {
  mipmap_comp proto;
  for (int d = 0; d < 6; d++) {
    proto.src_tex3D[d].levels = levels + (size_t)d * n_levels;
    proto.src_tex3D[d].R = R;
    proto.src_tex3D[d].n_levels = n_levels;
  }
  int current_dim = R, mip = 0;
  while (current_dim >= 1) {
    /* levels past the texture's storage cannot be bound: the reference's remaining dispatches have no effect */
    if (mip + 1 >= n_levels || (R >> (mip + 1)) < 1) break;
    proto.resolution = current_dim;
    proto.mip = mip;
    const int Nd = R >> (mip + 1);
    for (int d = 0; d < 6; d++) { proto.dest_tex3D[d].texels = levels[(size_t)d * n_levels + mip + 1]; proto.dest_tex3D[d].N = Nd; }
    /* glDispatchCompute(ceil(dim / 8))^3 groups of 8^3: invocations with an id >= resolution return at once (mipmap.comp:47-50);
     * those in [Nd, resolution) would fetch and store outside the images, which GL discards -- they are not run here */
#pragma omp parallel for schedule(static)
    for (int z = 0; z < Nd; z++) {
      mipmap_comp cs = proto;
      for (int y = 0; y < Nd; y++)
        for (int x = 0; x < Nd; x++) {
          cs.gl_GlobalInvocationID = uvec3(x, y, z);
          cs._init_globals();
          cs.main();
        }
    }
    mip++;
    current_dim /= 2;
  }
  return 0;
}

/* ---- vertex stage of Renderer::visualize() + the fixed-function camera pass ---- */
int run_gbuffer(const orc_scene_t* sc, const float view[16], const float proj[16], int W, int H, uint32_t* tri_id, float* depth,
                float* world_pos, float* normal, uint32_t* material) {
  using namespace vct_ff;
  cone_vert vs;
  vs.projection = load_mat4(proj);
  vs.view = load_mat4(view);
  CameraPass pass(W, H);
  uint32_t seq = 0;
  for (uint32_t d = 0; d < sc->n_draws; d++) {
    const orc_draw_t& dr = sc->draws[d];
    vs.model = load_mat4(dr.model);
    for (uint32_t t = 0; t + 3 <= dr.index_count; t += 3, seq++) {
      FFVertex in[3];
      for (int k = 0; k < 3; k++) {
        const orc_vertex_t& v = sc->verts[dr.vertex_base + sc->indices[dr.first_index + t + k]];
        vs.position = load3(v.pos); vs.normal = load3(v.norm); vs.uv = vec2(v.uv[0], v.uv[1]);
        vs._init_globals();
        vs.main();
        in[k].clip = V4{vs.gl_Position.x, vs.gl_Position.y, vs.gl_Position.z, vs.gl_Position.w};
        in[k].world = v3(vs.vs_out.world_position.x, vs.vs_out.world_position.y, vs.vs_out.world_position.z);
        in[k].nrm = v3(vs.vs_out.normal.x, vs.vs_out.normal.y, vs.vs_out.normal.z);
      }
      pass.add_triangle(in, dr.material, seq);
    }
  }
  pass.resolve(tri_id, depth, world_pos, normal, material);
  return 0;
}

/* ---- fragment stage of Renderer::visualize(), src/renderer.cpp:355-390, on a resolved G-buffer ---- */
void setup_cone_frag(cone_frag& fs, const orc_scene_t* sc, const float view[16], const uint32_t* const* levels, int R, int n_levels,
                     const orc_trace_params_t* prm) {
  fs.cube_size = sc->cube_size;   /* :365-366 */
  fs.cube_res = R;
  fs.enable_diffuse = prm->enable_diffuse != 0; fs.enable_specular = prm->enable_specular != 0;   /* :368-371 */
  fs.enable_shadow = prm->enable_shadow != 0; fs.enable_direct = prm->enable_direct != 0;
  fs.view_voxel_dir = prm->view_voxel_dir; fs.view_voxel_lod = prm->view_voxel_lod;               /* :373-374 */
  for (int d = 0; d < 6; d++) { fs.tex3D[d].levels = levels + (size_t)d * n_levels; fs.tex3D[d].R = R; fs.tex3D[d].n_levels = n_levels; }
  /* upload_camera :279-280: glm::column(view, 3) (sic) */
  const mat4 v = load_mat4(view);
  fs.camera_position = vec3(glm::column(v, 3));
  upload_lights(fs, sc);
}

int run_shade(const orc_scene_t* sc, const float view[16], int W, int H, const uint32_t* tri_id, const float* world_pos, const float* normal,
              const uint32_t* material, const uint32_t* const* levels, int R, int n_levels, const orc_trace_params_t* prm, int tile_stride,
              int tile_phase, uint32_t* frame) {
  using namespace vct_ff;
  cone_frag proto;
  setup_cone_frag(proto, sc, view, levels, R, n_levels, prm);
  if (tile_stride < 1) tile_stride = 1;
  const int tiles_x = (W + 31) / 32;
  const float unwritten = std::numeric_limits<float>::quiet_NaN();
#pragma omp parallel for schedule(dynamic, 1)
  for (int j = 0; j < H; j++) {
    cone_frag fs = proto;
    for (int i = 0; i < W; i++) {
      const int tile = (j / 32) * tiles_x + (i / 32);
      if (tile % tile_stride != tile_phase) continue;
      const size_t px = (size_t)j * W + i;
      if (tri_id[px] == 0xFFFFFFFFu) { frame[px] = kClearColour; continue; }
      bind_material(fs, sc->mats[material[px]]);
      fs.vs_out.world_position = vec4(world_pos[px * 3], world_pos[px * 3 + 1], world_pos[px * 3 + 2], 1.0f);
      fs.vs_out.normal = vec3(normal[px * 3], normal[px * 3 + 1], normal[px * 3 + 2]);
      fs.final_color = vec4(unwritten);
      fs._init_globals();
      fs.main();
      if (fs.final_color.x != fs.final_color.x && fs.final_color.w != fs.final_color.w) { frame[px] = kClearColour; continue; } /* returned before writing */
      float rgba[4] = {fs.final_color.x, fs.final_color.y, fs.final_color.z, fs.final_color.w};
      /* blend SRC_ALPHA / ONE_MINUS_SRC_ALPHA over the clear colour (:387-388); alpha = 1 for shaded pixels */
      if (prm->view_voxel_dir < 7) {
        const float bg[4] = {0.15f, 0.25f, 0.25f, 1.0f};
        const float a = rgba[3];
        for (int k = 0; k < 4; k++) rgba[k] = rgba[k] * a + bg[k] * (1.0f - a);
      }
      frame[px] = pack_unorm(rgba);
    }
  }
  return 0;
}

#if GLREF_RULES
/* ---- PERFORMANCE MODEL of the texture-unit work of the product's cone kernel (tools/tex_lane_model.py; not a parity check) ----
 * Every cone the reference's fragment shader would trace for the pixels of one 2x2 pixel block (= one TEX quad of cone_kernel_fast)
 * is marched with the shader's own trace_cone; a recorder in textureLod classifies every sample the way the kernel does
 * (csrc/cone_trace.cu trace_cone_fast): 0 = no texture fetch (footprint empty, or level 0 alone, which the kernel filters in
 * software), 1 = one level through the texture unit, 2 = two levels.  The TEX pipe works on whole quads, so a fetch costs the same
 * whether one or four lanes of a quad need it; the model adds up quad-level fetch units (one level of one direction of one quad) for
 *   [0] ideal: every lane's own need / 4 (perfectly packed quads)
 *   [1] the kernel as it is: a one-level and a two-level instruction are issued separately when the lanes of a quad disagree
 *   [2] quad-uniform decision: if any lane of the quad needs two levels all its fetching lanes take the two-level instruction
 *   [3] lane-autonomous march: every lane skips its own empty samples; the quad pays max-over-lanes per fetching round (+ rule [2])
 *   [4] samples (lane steps), [5] samples that fetch, [6] warp-level... (unused)
 * Units are multiplied by the number of directions with a non-zero weight (3 almost always). */
struct ConeSeq { std::vector<uint8_t> cls; };

inline void classify(const std::vector<TexRecord>& rec, int n_levels, ConeSeq& out) {
  out.cls.clear();
  for (size_t i = 0; i + 3 <= rec.size(); i += 3) {
    const float lod = rec[i].lod;
    const bool e0 = rec[i].empty_l0 && rec[i + 1].empty_l0 && rec[i + 2].empty_l0;
    const bool e1 = rec[i].empty_l1 && rec[i + 1].empty_l1 && rec[i + 2].empty_l1;
    uint8_t c;
    if (lod < 1.0f) c = (lod > 0.0f && !e1) ? 1 : 0;          /* level 1 through the texture unit, level 0 in software */
    else {
      const bool two = lod > floorf(lod);
      if (two) c = e1 ? 0 : (e0 ? 1 : 2);                       /* coarser level empty => finer empty too */
      else c = e0 ? 0 : 1;
    }
    out.cls.push_back(c);
  }
  (void)n_levels;
}

int run_tex_model(const orc_scene_t* sc, const float view[16], int W, int H, const uint32_t* tri_id, const float* world_pos, const float* normal,
                  const uint32_t* material, const uint32_t* const* levels, int R, int n_levels, const orc_trace_params_t* prm, int tile_stride,
                  int tile_phase, double out[8]) {
  cone_frag proto;
  setup_cone_frag(proto, sc, view, levels, R, n_levels, prm);
  if (tile_stride < 1) tile_stride = 1;
  const int tiles_x = (W + 31) / 32;
  double t_ideal = 0, t_cur = 0, t_quad = 0, t_auto = 0, t_samples = 0, t_fetching = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : t_ideal, t_cur, t_quad, t_auto, t_samples, t_fetching)
  for (int qy = 0; qy < H / 2; qy++) {
    cone_frag fs = proto;
    std::vector<TexRecord> rec;
    std::vector<ConeSeq> seq[4];
    for (int qx = 0; qx < W / 2; qx++) {
      const int tile = ((2 * qy) / 32) * tiles_x + ((2 * qx) / 32);
      if (tile % tile_stride != tile_phase) continue;
      size_t n_jobs = 0;
      for (int l = 0; l < 4; l++) {
        seq[l].clear();
        const int x = 2 * qx + (l & 1), y = 2 * qy + (l >> 1);
        const size_t px = (size_t)y * W + x;
        if (tri_id[px] == 0xFFFFFFFFu) continue;
        bind_material(fs, sc->mats[material[px]]);
        fs.vs_out.world_position = vec4(world_pos[px * 3], world_pos[px * 3 + 1], world_pos[px * 3 + 2], 1.0f);
        fs.vs_out.normal = vec3(normal[px * 3], normal[px * 3 + 1], normal[px * 3 + 2]);
        fs._init_globals();
        const vec3 pos = fs.scale_and_bias(vec3(fs.vs_out.world_position.x, fs.vs_out.world_position.y, fs.vs_out.world_position.z) / fs.cube_size);
        if (!fs.within_cube(pos, 0)) continue;
        const vec3 n = fs.vs_out.normal;
        const vec3 view_dir = normalize(vec3(fs.vs_out.world_position.x, fs.vs_out.world_position.y, fs.vs_out.world_position.z) - fs.camera_position);
        /* the cones of main(): voxel_cone_tracing.frag:140-168 (diffuse), :175-218 (shadow), :229-241 (specular, refraction) */
        struct Cone { vec3 dir; float ap, md; };
        std::vector<Cone> cones;
        const float TAN = 0.55785173935f, MAXD = 1.73205080757f;
        if (fs.enable_diffuse) {
          const vec3 o1 = normalize(fs.tangent(n)), o2 = normalize(cross(o1, n));
          const vec3 c1 = 0.5f * (o1 + o2), c2 = 0.5f * (o1 - o2);
          const vec3 d[9] = {n, mix(n, o1, 0.5f), mix(n, -o1, 0.5f), mix(n, o2, 0.5f), mix(n, -o2, 0.5f), mix(n, c1, 0.5f), mix(n, -c1, 0.5f), mix(n, c2, 0.5f), mix(n, -c2, 0.5f)};
          for (int i = 0; i < 9; i++) cones.push_back(Cone{d[i], TAN, MAXD});
        }
        if (fs.enable_direct && fs.enable_shadow)
          for (int i = 0; i < fs.point_light_count && i < 10; i++) {
            const vec3 lp = fs.scale_and_bias(fs.point_lights[i].position / fs.cube_size);
            vec3 ld = lp - pos;
            const float d = length(ld);
            cones.push_back(Cone{normalize(ld), 0.1f, d});
          }
        if (fs.enable_specular) {
          cones.push_back(Cone{normalize(reflect(-view_dir, n)), fs.specular_aperture, MAXD});
          if (fs.illum == 4 || fs.illum == 6 || fs.illum == 7 || fs.illum == 9) cones.push_back(Cone{refract(view_dir, n, 1.0f / fs.ior), fs.specular_aperture, MAXD});
        }
        seq[l].resize(cones.size());
        for (size_t c = 0; c < cones.size(); c++) {
          rec.clear();
          tex_recorder() = &rec;
          (void)fs.trace_cone(pos, cones[c].dir, cones[c].ap, cones[c].md);
          tex_recorder() = nullptr;
          classify(rec, n_levels, seq[l][c]);
        }
        n_jobs = std::max(n_jobs, cones.size());
      }
      for (size_t j = 0; j < n_jobs; j++) {
        const std::vector<uint8_t>* s[4];
        static const std::vector<uint8_t> none;
        size_t steps = 0;
        for (int l = 0; l < 4; l++) { s[l] = j < seq[l].size() ? &seq[l][j].cls : &none; steps = std::max(steps, s[l]->size()); }
        std::vector<uint8_t> packed[4];
        for (size_t k = 0; k < steps; k++) {
          int any1 = 0, any2 = 0;
          for (int l = 0; l < 4; l++) {
            if (k >= s[l]->size()) continue;
            const uint8_t c = (*s[l])[k];
            t_samples += 1; t_ideal += c / 4.0;
            if (c) { t_fetching += 1; packed[l].push_back(c); }
            any1 |= c == 1; any2 |= c == 2;
          }
          t_cur += any1 * 1 + any2 * 2;
          t_quad += any2 ? 2 : any1;
        }
        size_t rounds = 0;
        for (int l = 0; l < 4; l++) rounds = std::max(rounds, packed[l].size());
        for (size_t k = 0; k < rounds; k++) {
          int m = 0;
          for (int l = 0; l < 4; l++) if (k < packed[l].size()) m = std::max<int>(m, packed[l][k]);
          t_auto += m;
        }
      }
    }
  }
  out[0] = 3 * t_ideal; out[1] = 3 * t_cur; out[2] = 3 * t_quad; out[3] = 3 * t_auto; out[4] = t_samples; out[5] = t_fetching; out[6] = out[7] = 0;
  return 0;
}
#endif

}  // namespace GLREF_NS

/* ================================================================== */
extern "C" {

/* tex[0..5]: the six level-0 images (R^3 uint32 each), cleared by the callee like clear_tex_3d */
int GLREF_FN(voxelize)(const orc_scene_t* sc, int R, uint32_t* const* tex, uint64_t* n_fragments) {
  if (!sc || !tex || R <= 0) return -1;
  return GLREF_NS::run_voxelize(sc, R, tex, n_fragments, 0, 0, 0u);
}

/* sensitivity studies only: another fragment order / barycentric rule (see run_voxelize) */
int GLREF_FN(voxelize_variant)(const orc_scene_t* sc, int R, uint32_t* const* tex, uint64_t* n_fragments, int order, int bary, uint32_t seed) {
  if (!sc || !tex || R <= 0 || order < 0 || order > 3 || bary < 0 || bary > 1) return -1;
  return GLREF_NS::run_voxelize(sc, R, tex, n_fragments, order, bary, seed);
}

/* levels[d * n_levels + l]: level 0 of every direction filled by the caller */
int GLREF_FN(mipmap)(uint32_t* const* levels, int R, int n_levels) {
  if (!levels || R <= 0 || n_levels < 1) return -1;
  return GLREF_NS::run_mipmap(levels, R, n_levels);
}

int GLREF_FN(gbuffer)(const orc_scene_t* sc, const float view[16], const float proj[16], int W, int H, uint32_t* tri_id, float* depth,
                      float* world_pos, float* normal, uint32_t* material) {
  if (!sc || !tri_id || !depth) return -1;
  return GLREF_NS::run_gbuffer(sc, view, proj, W, H, tri_id, depth, world_pos, normal, material);
}

int GLREF_FN(shade)(const orc_scene_t* sc, const float view[16], int W, int H, const uint32_t* tri_id, const float* world_pos,
                    const float* normal, const uint32_t* material, const uint32_t* const* levels, int R, int n_levels,
                    const orc_trace_params_t* prm, int tile_stride, int tile_phase, uint32_t* frame) {
  if (!sc || !tri_id || !world_pos || !normal || !material || !levels || !prm || !frame) return -1;
  return GLREF_NS::run_shade(sc, view, W, H, tri_id, world_pos, normal, material, levels, R, n_levels, prm, tile_stride, tile_phase, frame);
}

#if GLREF_RULES
/* performance model of the cone kernel's texture-unit work (see run_tex_model); out[0..5] */
int glref_rules_tex_model(const orc_scene_t* sc, const float view[16], int W, int H, const uint32_t* tri_id, const float* world_pos,
                          const float* normal, const uint32_t* material, const uint32_t* const* levels, int R, int n_levels,
                          const orc_trace_params_t* prm, int tile_stride, int tile_phase, double* out) {
  if (!sc || !tri_id || !world_pos || !normal || !material || !levels || !prm || !out) return -1;
  return GLREF_NS::run_tex_model(sc, view, W, H, tri_id, world_pos, normal, material, levels, R, n_levels, prm, tile_stride, tile_phase, out);
}
#endif

#if !GLREF_RULES
/* ---- the reference's HOST-SIDE matrix code, to pin the product's camera / model-matrix helpers (scene.py, include/vct/math.h) ----
 * Camera is the reference's own struct, compiled from src/camera.h where it lies (GLM only); the model matrix follows
 * src/main.cpp:369-372 with the identity start the author intended (glm 0.9.9 leaves `glm::mat4 m;` uninitialised). */
}  /* extern "C" */
#include <glm/gtc/matrix_transform.hpp>
#include "camera.h"
extern "C" {
int glref_camera(const float eye[3], float pitch_deg, float yaw_deg, float lens_angle, float aspect, float z_near, float z_far, float view[16],
                 float proj[16]) {
  Camera cam(glm::vec3(eye[0], eye[1], eye[2]), pitch_deg, yaw_deg);       /* main.cpp:107 */
  cam.set_perspective(lens_angle, aspect, z_near, z_far);                     /* main.cpp:108: 45.0f, taken as radians by glm 0.9.9 */
  const glm::mat4 v = cam.get_lookat(), p = cam.get_projection();             /* main.cpp:109 */
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 4; r++) { view[c * 4 + r] = v[c][r]; proj[c * 4 + r] = p[c][r]; }
  return 0;
}
int glref_model_trs(const float t[3], float rot_y, float scale, float out[16]) {
  glm::mat4 m(1.0f);
  m = glm::translate(m, glm::vec3(t[0], t[1], t[2]));                         /* main.cpp:370 */
  m = glm::rotate(m, rot_y, glm::vec3(0, 1, 0));                              /* main.cpp:371 */
  m = glm::scale(m, glm::vec3(scale, scale, scale));                          /* main.cpp:372 */
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 4; r++) out[c * 4 + r] = m[c][r];
  return 0;
}
#endif

/* trace_cone() of voxel_cone_tracing.frag:88-119 on its own (float results, no 8-bit rounding in between) */
int GLREF_FN(trace_cone)(const uint32_t* const* levels, int R, int n_levels, const float origin[3], const float dir[3], float aperture,
                         float max_dist, float out_rgba[4]) {
  using namespace GLREF_NS;
  cone_frag fs;
  for (int d = 0; d < 6; d++) { fs.tex3D[d].levels = levels + (size_t)d * n_levels; fs.tex3D[d].R = R; fs.tex3D[d].n_levels = n_levels; }
  fs.cube_res = R;
  fs.cube_size = 1.0f;
  fs.shininess = 1.0f;
  fs._init_globals();
  const vec4 r = fs.trace_cone(load3(origin), load3(dir), aperture, max_dist);
  out_rgba[0] = r.x; out_rgba[1] = r.y; out_rgba[2] = r.z; out_rgba[3] = r.w;
  return 0;
}

/* imageAtomicRGBA8Avg of voxelize.frag:95-120 applied to one texel that holds `stored` */
uint32_t GLREF_FN(rgba8_avg)(uint32_t stored, const float val01[4]) {
  using namespace GLREF_NS;
  voxelize_frag fs;
  uint32_t texel = stored;
  uimage3D img; img.texels = &texel; img.N = 1;
  fs.imageAtomicRGBA8Avg(img, ivec3(0), vec4(val01[0], val01[1], val01[2], val01[3]));
  return texel;
}

/* axis selection of voxelize.geom:25-55: 0 = (x,y), 1 = (y,z), 2 = (x,z) projection, recognised from the emitted positions */
int GLREF_FN(select_axis)(const float wp0[3], const float wp1[3], const float wp2[3]) {
  using namespace GLREF_NS;
  GeomRun gs;
  const float* w[3] = {wp0, wp1, wp2};
  for (int k = 0; k < 3; k++) { gs.vs_out[k].world_position = vec4(load3(w[k]), 1.0f); gs.vs_out[k].normal = vec3(0.0f); gs.vs_out[k].uv = vec2(0.0f); }
  gs._init_globals();
  gs.main();
  if (gs.emitted.size() != 3) return -1;
  /* which pair of coordinates was copied into gl_Position.xy for all three vertices? */
  const int pairs[3][2] = {{0, 1}, {1, 2}, {0, 2}};
  int found = -1, n_found = 0;
  for (int a = 0; a < 3; a++) {
    bool ok = true;
    for (int k = 0; k < 3; k++) ok = ok && gs.emitted[k].pos.x == w[k][pairs[a][0]] && gs.emitted[k].pos.y == w[k][pairs[a][1]];
    if (ok) { found = a; n_found++; }
  }
  return n_found == 1 ? found : -2;   /* -2: ambiguous input (choose vertices with distinct coordinates) */
}

}  /* extern "C" */
