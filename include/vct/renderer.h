// renderer.h -- the reference's `struct Renderer` (src/renderer.h:122-196) on the B200-native path.
//
// Same public method names, argument meaning and error behaviour as the reference:
//   * ids are indices; failures print to stderr and return INVALID_ID (load_model, renderer.cpp:417-421);
//     queue_model with a bad id is silently ignored (renderer.cpp:211-212); nothing throws.
//   * render() = camera update -> voxelize() [clear, scatter, filter()] -> visualize(), then the per-frame
//     queues are cleared (renderer.cpp:392-405).
// What changed underneath: no OpenGL.  The device layer is vct::Device (CUDA), the six GL 3-D textures are
// one vct_grid_t, the three GLSL programs are compiled-in sm_100a kernels (load_shader() keeps its
// signature and returns the fixed id of the built-in program of that name), UBOs are plain device arrays,
// and the frame lands in a device RGBA8 buffer that read_frame() copies out (the reference's default
// framebuffer + glfwSwapBuffers).  Every GPU call goes through the C ABI of vct_c.h; there is no CPU
// fallback: when no sm_100 device is usable ok() is false and render() does nothing but report it.
//
// glm types: any 64-byte column-major matrix (glm::mat4, vct::mat4) and any 12-byte vec3 are accepted.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

#include "vct/device.h"
#include "vct/math.h"
#include "vct/texture_3d.h"

#define MAX_TEXTURE_BINDINGS 32
#ifndef INVALID_ID
#define INVALID_ID (~(size_t)0)
#endif

namespace vct {

typedef size_t model_id_t;
typedef size_t shader_id_t;
typedef size_t material_id_t;

typedef struct { size_t start; size_t size; } buffer_range_t;         // src/renderer.h:30-34

// src/renderer.h:37-42; `ubo` is the slot of the material in the device material array
typedef struct { unsigned ubo; unsigned textures[MAX_TEXTURE_BINDINGS]; size_t num_textures; } material_t;

typedef struct { buffer_range_t range; unsigned draw_type; material_id_t material_id; } draw_obj_t;  // :57-64

typedef struct {                                                        // :67-79 (vao/vbo/ebo -> buffer offsets)
  size_t vertex_base, index_base;
  buffer_range_t draw_obj_range;
  shader_id_t shader_id;
  vec3 dimensions;
  mat4 model_matrix;
} model_t;

typedef struct { mat4 projection; mat4 view; } camera_data_t;          // :81-85
typedef struct { vec3 position; vec3 color; float intensity; } point_light_t;  // :87-92, 28 bytes

#pragma pack(push, 1)
typedef struct {                                                        // :94-120, 128 bytes, std140-compatible
  vec4 ambient, diffuse, specular, transmittance;
  vec3 emission;
  float shininess, ior, dissolve;
  int illum;
  float roughness, metallic, sheen, clearcoat_thickness, clearcoat_roughness, anisotropy, anisotropy_rotation;
  float pad[2];
} material_data_t;
#pragma pack(pop)
static_assert(sizeof(material_data_t) == 128 && sizeof(point_light_t) == 28, "layouts of src/renderer.h:87-120");

struct Renderer {
  Renderer(int width, int height, int cuda_ordinal = 0);
  ~Renderer();
  Renderer(const Renderer&) = delete;
  Renderer& operator=(const Renderer&) = delete;

  // Resource
  model_id_t load_model(const char* filename);
  shader_id_t load_shader(const char* vertex_shader_name, const char* fragment_shader_name, const char* geometry_shader_name = NULL);
  material_id_t add_material(material_data_t& material_data);
  void upload_material_data(material_data_t& material_data, material_id_t material_id);
  void set_model_material(material_id_t material_id, model_id_t model_id);

  vec3 get_model_dimensions(model_id_t model);
  void set_grid_resolution(unsigned int res);
  void set_grid_size(float size);

  template <class M> void set_camera_transform(M& lookat, M& projection) {
    static_assert(sizeof(M) == 64, "column-major 4x4 float matrix expected");
    std::memcpy(m_camera.view.m, &lookat, 64);
    std::memcpy(m_camera.projection.m, &projection, 64);
  }
  template <class M> void set_model_transform(model_id_t model_id, M model_matrix) {
    static_assert(sizeof(M) == 64, "column-major 4x4 float matrix expected");
    if (model_id < m_models.size()) std::memcpy(m_models[model_id].model_matrix.m, &model_matrix, 64);
  }
  material_t& get_material(material_id_t material_id);

  // Render
  void set_rendering_phases(bool direct, bool diffuse, bool specular, bool shadow);
  void set_voxel_view_dir(int dir, float lod);
  void queue_model(model_id_t model_id);
  void queue_point_light(point_light_t& point_light);
  void render();

  // ---- additions of the CUDA build (no reference counterpart) ----
  bool ok() const { return m_device.ok() && m_scene && m_grid && m_target; }
  Device& device() { return m_device; }
  int width() const { return m_viewport_width; }
  int height() const { return m_viewport_height; }
  bool read_frame(uint32_t* rgba8);                     // W*H RGBA8, row 0 = bottom (GL window coordinates); synchronises
  uint64_t read_frame_async(uint32_t* pinned_rgba8);    // same, on a copy stream that overlaps the next render(); 0 on failure, else a ticket
  bool wait_frame(uint64_t ticket);                     // blocks until that read-back has landed
  void* frame_device_ptr();                             // device pointer of the RGBA8 frame
  bool read_voxels(int level, int dir, uint32_t* rgba8); // one level of one directional texture (glGetTexImage)
  void set_sampler(int vct_sampler) { m_sampler = vct_sampler; }  // VCT_SAMPLER_FP32 / VCT_SAMPLER_TEX
  // storage of the voxel pyramid: VCT_GRID_RGBA8 with 7 levels is the reference (texture_3d.cpp:3-25, renderer.cpp:186); VCT_GRID_RGBA16F and/or
  // another level count (0 = reference default for RGBA8, the full chain for RGBA16F) is BASELINE config 5's variant.  Recreates the grid.
  void set_grid_storage(int vct_grid_format, int levels = 0);
  // VCT_ACCUM_ORDERED (the reference's running average, bit-exact, default) / VCT_ACCUM_FIXED_POINT (order-independent integer mean)
  bool set_voxel_accumulation(int vct_accum_mode) { return vct_voxelize_set_accum_mode(m_device.handle(), vct_accum_mode) == VCT_OK; }
  void set_diffuse_cone_count(int n) { m_diffuse_cones = n; }     // 9 = reference, 5 = BASELINE.json variant
  void set_rank(int rank, int nranks) { m_rank = rank; m_nranks = nranks; }  // multi-GPU: z-slab / screen-tile share of this process
  void set_staged_passes(bool on) { m_staged = on; }   // render() = voxelize(); visualize() as separate stage calls instead of one vct_render_frame
  int last_error() const { return m_last_rc; }         // VCT_OK or the error code of the last render()
  const material_data_t& get_material_data(material_id_t id) const { return m_material_data[id]; }
  size_t triangle_count() const { return m_indices.size() / 3; }

 private:
  void draw_models();   // flattens m_draw_queue into the device draw list (renderer.cpp:240-257)
  void upload_lights(); // renderer.cpp:259-273
  // The reference's private passes, kept as the stage-by-stage path (set_staged_passes(true): each is one C-ABI stage call, the
  // G-buffer pass runs in line).  render() otherwise issues ONE vct_render_frame, which overlaps the G-buffer pass with
  // clear + voxelize + filter on a second stream.  (upload_camera, renderer.cpp:275-281, has no counterpart: the matrices are
  // kernel arguments, there is no UBO.)
  bool voxelize();      // renderer.cpp:316-353: clear + scatter, then filter()
  bool filter();        // renderer.cpp:283-314
  bool visualize();     // renderer.cpp:355-390
  vct_trace_params_t trace_params() const;
  bool upload_geometry();
  bool upload_materials();

  Device m_device;
  vct_scene_t* m_scene = nullptr;
  vct_grid_t* m_grid = nullptr;
  vct_target_t* m_target = nullptr;

  // Lifetime
  std::vector<material_t> m_materials;
  std::vector<material_data_t> m_material_data;
  std::vector<model_t> m_models;
  std::vector<draw_obj_t> m_draw_objs;
  std::vector<vct_vertex_t> m_vertices;   // all models, concatenated (model_t::vertex_base)
  std::vector<uint32_t> m_indices;        // all models, concatenated (model_t::index_base)
  bool m_geometry_dirty = false, m_materials_dirty = false;

  camera_data_t m_camera;
  float m_cube_size = 1.0f;
  size_t m_resolution = 0;
  int m_viewport_width, m_viewport_height;
  bool m_enable_shadows = true, m_enable_direct = true, m_enable_indirect_diffuse = true, m_enable_indirect_specular = true;
  int m_view_voxel_dir = 7;
  float m_view_voxel_lod = 0.0f;
  int m_sampler = VCT_SAMPLER_TEX, m_diffuse_cones = 9, m_rank = 0, m_nranks = 1;
  int m_grid_format = VCT_GRID_RGBA8, m_grid_levels = 0;
  static constexpr int VCT_MAX_LEVELS_HOST = 12;
  bool m_staged = false;
  int m_last_rc = 0;

  // Per frame
  std::vector<model_t> m_draw_queue;
  std::vector<point_light_t> m_point_lights;

  shader_id_t m_voxelize_shader = 0, m_draw_shader = 1, m_mipmap_shader = 2;
};

// OBJ/MTL reader used by load_model (tinyobjloader v1.1.0 semantics as the reference relies on them,
// thirdparty/tinyobjloader/tiny_obj_loader.h; vertex dedupe + per-material ranges of renderer.cpp:460-556)
struct loaded_mesh_t {
  std::vector<vct_vertex_t> vertices;
  std::vector<uint32_t> indices;
  struct range_t { uint32_t first_index, index_count; int material; };   // material = index into `materials`, -1 = none
  std::vector<range_t> ranges;
  std::vector<material_data_t> materials;
  vec3 bbox_min, bbox_max;
};
bool load_obj(const char* filename, loaded_mesh_t* out);

}  // namespace vct

#ifndef VCT_NO_GLOBAL_NAMES   // the reference declares these at global scope
using vct::Renderer;
using vct::model_id_t;
using vct::shader_id_t;
using vct::material_id_t;
using vct::material_t;
using vct::material_data_t;
using vct::point_light_t;
#endif
