/*
 * vct_oracle.h -- C interface of the CPU ORACLE for the voxel-cone-tracing hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * PARITY: the PROGRAMMABLE stages are pinned to the reference itself -- its six GLSL shaders, read from /root/reference where
 * they lie, are rewritten syntactically, compiled against the reference's vendored GLM and executed on the CPU
 * (oracle/glsl_ref/ -> oracle/_ref/libvct_glsl_ref.so); voxel grid, all mip volumes, G-buffer attributes and the frame of this
 * oracle equal that program's BIT FOR BIT (tests/test_glsl_ref.py; vectors it produced: tests/golden/glsl_ref_vectors.json,
 * tests/test_glsl_ref_golden.py).  The FIXED-FUNCTION stages are held against a real OpenGL implementation, not the one BASELINE.json
 * names (no OpenGL 4.5 driver exists in this image) but the Mesa 18.1 llvmpipe inside Nsight Compute, driven through all three passes of
 * the reference with its own shader text (oracle/gl_ref/, tests/test_gl_llvmpipe.py): the fragments of the voxelization pass (count,
 * occupancy and sample counts identical), the mip chain (bit for bit once two liberties of that driver are modelled, one LSB on rounding
 * ties otherwise), raster coverage / clipping / interpolation / depth test / textureLod / blending / unorm conversion of the camera pass
 * and whole frames (within 1/255).  What stays a WRITTEN RULE with no GL behind it: R4, the order in which fragments reach the running
 * average (GL defines none; llvmpipe 18.1 has no image atomics to show one), and -- llvmpipe filtering mip levels "brilinearly" -- the
 * linear LOD blend of R7 at fractions other than 0 and 0.5 (the specification's formula).  Beside that: known-answer tests derived by
 * hand from the shader text (tests/test_oracle_kat.py) and the tinyobjloader cross-check of the scene inputs.
 */
#ifndef VCT_ORACLE_H
#define VCT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 32-byte vertex, same packing as the reference's vert_data_t (src/renderer.cpp:24-35). */
typedef struct { float pos[3]; float norm[3]; float uv[2]; } orc_vertex_t;

/* One glDrawElements range (reference draw_obj_t src/renderer.h:57-64 + model matrix
 * src/renderer.h:78).  model is column-major like glm::mat4. */
typedef struct {
  uint32_t first_index, index_count, vertex_base, material;
  float model[16];
} orc_draw_t;

/* reference point_light_t src/renderer.h:87-92 (28 bytes) */
typedef struct { float position[3]; float color[3]; float intensity; } orc_light_t;

/* reference material_data_t src/renderer.h:94-120 (128 bytes, pack(1)) */
typedef struct {
  float ambient[4], diffuse[4], specular[4], transmittance[4];
  float emission[3];
  float shininess, ior, dissolve;
  int32_t illum;
  float roughness, metallic, sheen, clearcoat_thickness, clearcoat_roughness, anisotropy,
      anisotropy_rotation;
  float pad[2];
} orc_material_t;

typedef struct {
  const orc_vertex_t* verts;   uint32_t n_verts;
  const uint32_t* indices;     uint32_t n_indices;
  const orc_draw_t* draws;     uint32_t n_draws;
  const orc_material_t* mats;  uint32_t n_mats;
  const orc_light_t* lights;   uint32_t n_lights;
  float cube_size;
} orc_scene_t;

/* statistics the tests / bench use */
typedef struct {
  uint64_t fragments;        /* voxelization fragments that passed the bounds check   */
  uint64_t fragments_oob;    /* fragments whose voxel was outside the grid            */
  uint64_t occupied;         /* voxels != 0 after voxelization                        */
  uint64_t max_per_voxel;    /* max fragments folded into one voxel                   */
  uint64_t wrapped_voxels;   /* voxels that received >= 16 fragments (count wrap)     */
  uint64_t tris_no_frag;     /* triangles that produced no fragment                   */
} orc_voxel_stats_t;

typedef struct {
  uint64_t shaded_pixels;
  uint64_t samples_diffuse, samples_shadow, samples_specular, samples_refraction;
} orc_trace_stats_t;

typedef struct {
  int enable_direct, enable_diffuse, enable_specular, enable_shadow;
  int view_voxel_dir;      /* >= 7: normal shading (reference main.cpp:300) */
  float view_voxel_lod;
  int n_diffuse_cones;     /* 9 = reference; 5 = {normal + 4 side}/5 variant */
} orc_trace_params_t;

/* ---- pieces with known answers ---- */
uint32_t orc_rgba8_avg_fold(uint32_t stored, const float val01[4]);   /* voxelize.frag:95-120, sequential */
int orc_select_axis(const float wp0[3], const float wp1[3], const float wp2[3]); /* voxelize.geom:25-55: 0=(x,y) 1=(y,z) 2=(x,z) */
void orc_perspective(float fovy, float aspect, float zn, float zf, float out[16]);   /* glm::perspective RH, -1..1 */
void orc_look_at(const float eye[3], const float center[3], const float up[3], float out[16]);
void orc_camera_front(float pitch_deg, float yaw_deg, float out[3]);  /* camera.h:25-37 */
float orc_specular_aperture(float shininess);                         /* voxel_cone_tracing.frag:220-227 */
/* one cone through a pyramid; returns number of loop iterations */
int orc_trace_cone(const uint32_t* const* levels /*[6*n_levels] dir-major: levels[d*n_levels+l]*/, int R, int n_levels,
                   const float origin[3], const float dir[3], float aperture, float max_dist, float out_rgba[4]);
void orc_texture_lod(const uint32_t* const* levels, int R, int n_levels, int dir, const float pos[3], float lod, float out[4]);

/* ---- stages ---- */
/* base: R^3 u32 (z-major: [z][y][x]); zeroed by the callee (clear_tex_3d) */
int orc_voxelize(const orc_scene_t* scene, int R, uint32_t* base, orc_voxel_stats_t* stats);
/* same, restricted to voxel z in [z0,z1) (multi-GPU slab semantics); base is still R^3 */
int orc_voxelize_slab(const orc_scene_t* scene, int R, int z0, int z1, uint32_t* base, orc_voxel_stats_t* stats);
/* accum_mode 0 = the reference's ordered running average; 1 = the non-reference fixed-point variant (rounded integer mean, order independent) */
int orc_voxelize_slab_mode(const orc_scene_t* scene, int R, int z0, int z1, int accum_mode, uint32_t* base, orc_voxel_stats_t* stats);
/* ---- RGBA16F storage variant (BASELINE.json config 5: "fp16 RGBA + full mip chain"; NOT the reference's format): texels are uint64 = four
 *      halves.  accum_mode 2 of orc_voxelize_slab_mode writes such a grid (mean colour in [0,1], `base` = R^3 uint64); fmt = 1 below. ---- */
int orc_mipmap_fmt(const uint32_t* base, int R, int n_levels, uint32_t* const* out, int fmt);
int orc_trace_cone_fmt(const uint32_t* const* levels, int R, int n_levels, int fmt, const float origin[3], const float dir[3], float aperture,
                       float max_dist, float out_rgba[4]);
int orc_trace_fmt(const orc_scene_t* scene, const float view[16], int W, int H, const uint32_t* tri_id, const float* world_pos, const float* normal,
                  const uint32_t* material, const uint32_t* const* levels, int R, int n_levels, int fmt, const orc_trace_params_t* prm, int row0, int row1,
                  int tile_stride, int tile_phase, uint32_t* frame, orc_trace_stats_t* stats);

/* pyramid in the reference's layout: 6 textures x n_levels; level l of direction d is
 * out[d*n_levels + l] with (R>>l)^3 texels.  Level 0 of every direction is filled with a
 * copy of base by this call (the reference writes the same value to all six, voxelize.frag:159-160) */
int orc_mipmap(const uint32_t* base, int R, int n_levels, uint32_t* const* out);

/* visibility: per pixel the global triangle sequence number of the GL_LESS winner
 * (0xFFFFFFFF = background) and its window depth; row 0 = bottom (GL) */
int orc_gbuffer(const orc_scene_t* scene, const float view[16], const float proj[16], int W, int H,
                uint32_t* tri_id, float* depth, float* world_pos /*3/px*/, float* normal /*3/px*/, uint32_t* material);

/* shade; frame is RGBA8, row 0 = bottom (GL window coordinates) */
int orc_trace(const orc_scene_t* scene, const float view[16], int W, int H,
              const uint32_t* tri_id, const float* world_pos, const float* normal, const uint32_t* material,
              const uint32_t* const* levels, int R, int n_levels, const orc_trace_params_t* params,
              int row0, int row1, int tile_stride, int tile_phase,
              uint32_t* frame, orc_trace_stats_t* stats);

int orc_num_threads(void);
/* TEST-ONLY: Renderer::visualize as a forward renderer (every depth-passing fragment shaded and blended in draw order); see vct_oracle.cpp */
int orc_render_forward(const orc_scene_t* scene, const float view[16], const float proj[16], int W, int H, const uint32_t* const* levels, int R,
                       int n_levels, const orc_trace_params_t* params, uint32_t* frame);
void orc_set_num_threads(int n);
/* TEST SWITCH: 0 = rule R7 (default), 1 = Mesa llvmpipe's brilinear mip filter (vct_fixed_function.h lod_filter_mode); used only by
 * tests/test_gl_llvmpipe.py to compare the oracle with the reference's shaders running on that driver. */
void orc_debug_set_lod_filter(int mode);
/* TEST SWITCH: 0 = rule R6 (unorm8 -> float = c / 255), 1 = c * (1.0f / 255.0f) as Mesa llvmpipe converts texels; tests/test_gl_llvmpipe.py only. */
void orc_debug_set_unorm_unpack(int mode);
/* TEST SWITCH: 1 = the four alpha_blend terms of the mip filter added as (t0 + t1) + (t2 + t3) -- Mesa's GLSL compiler rebalances the sum
 * mipmap.comp writes left to right (GLSL fixes no evaluation order without `precise`); tests/test_gl_llvmpipe.py only. */
void orc_debug_set_mip_balanced_sum(int on);

#ifdef __cplusplus
}
#endif
#endif
