// renderer.cpp -- vct::Renderer: the reference's frame orchestration (src/renderer.cpp:100-677) driving
// the CUDA hot path through the C ABI.  Host logic only: scene bookkeeping, draw-list flattening, one
// stream-ordered frame sequence per render().  No OpenGL, no compute, no CPU fallback.
#include "vct/renderer.h"

#include <cstdio>
#include <string>

namespace vct {

#define DEFAULT_MATERIAL_INDEX 0   // src/renderer.cpp:83

// fill_default_mat_data (src/renderer.cpp:85-98).  The reference then uploads the wrong struct
// (renderer.cpp:120-126, host-side bug listed in SURVEY 8a "UB to avoid copying"); the intent is kept.
static material_data_t default_mat_data() {
  material_data_t m;
  std::memset(&m, 0, sizeof m);
  m.ambient = {1, 0, 1, 0};
  m.diffuse = {1, 0, 1, 0};
  m.shininess = 1; m.ior = 1; m.dissolve = 1; m.illum = 0;
  return m;
}

static bool check(int rc, const char* what) {
  if (rc == VCT_OK) return true;
  std::fprintf(stderr, "vct::Renderer: %s failed (%d): %s\n", what, rc, vct_last_error());
  return false;
}

Renderer::Renderer(int width, int height, int cuda_ordinal) : m_device(cuda_ordinal), m_viewport_width(width), m_viewport_height(height) {
  m_camera.view = mat4::identity();
  m_camera.projection = mat4::identity();
  // default material = id 0 (renderer.cpp:115-128)
  material_data_t def = default_mat_data();
  add_material(def);
  // the three programs of renderer.cpp:131-141 are compiled-in kernels; the calls are kept for their ids
  m_draw_shader = load_shader("shader/voxel_cone_tracing.vert", "shader/voxel_cone_tracing.frag");
  m_voxelize_shader = load_shader("shader/voxelize.vert", "shader/voxelize.frag", "shader/voxelize.geom");
  if (!m_device.ok()) return;
  if (!check(vct_scene_create(m_device.handle(), &m_scene), "vct_scene_create")) return;
  if (!check(vct_target_create(m_device.handle(), width, height, &m_target), "vct_target_create")) return;
  set_grid_resolution(128);   // renderer.cpp:146
}

Renderer::~Renderer() {
  // renderer.cpp:149-165: everything the Renderer created dies with it
  if (m_grid) vct_grid_destroy(m_grid);
  if (m_target) vct_target_destroy(m_target);
  if (m_scene) vct_scene_destroy(m_scene);
}

shader_id_t Renderer::load_shader(const char* vertex_shader_name, const char* /*fragment*/, const char* geometry_shader_name) {
  // GLSL sources are not compiled any more; the id tells which built-in pass the name maps to
  if (geometry_shader_name) return 0;                                                             // voxelize pass
  if (vertex_shader_name && std::string(vertex_shader_name).find("voxel_cone_tracing") != std::string::npos) return 1;  // draw pass
  return 2;
}

model_id_t Renderer::load_model(const char* filename) {
  loaded_mesh_t mesh;
  if (!filename || !load_obj(filename, &mesh)) {
    std::fprintf(stderr, "Error loading file %s\n", filename ? filename : "(null)");   // renderer.cpp:417-421
    return INVALID_ID;
  }
  const size_t material_base_index = m_materials.size();   // renderer.cpp:424-430
  for (material_data_t& m : mesh.materials) add_material(m);

  model_t model;
  std::memset(&model, 0, sizeof model);
  model.vertex_base = m_vertices.size();
  model.index_base = m_indices.size();
  model.draw_obj_range.start = m_draw_objs.size();
  model.draw_obj_range.size = mesh.ranges.size();
  model.shader_id = m_draw_shader;
  model.model_matrix = mat4::identity();   // glm::mat4 is uninitialised in the reference (main.cpp:369,376); identity is the intent
  model.dimensions = mesh.bbox_max - mesh.bbox_min;   // renderer.cpp:517 (the min/max typos of :455-487 are not reproduced)
  for (const loaded_mesh_t::range_t& r : mesh.ranges) {
    draw_obj_t d;
    d.range.start = r.first_index;
    d.range.size = r.index_count;
    d.draw_type = 4;   // GL_TRIANGLES
    d.material_id = r.material < 0 ? DEFAULT_MATERIAL_INDEX : material_base_index + (size_t)r.material;   // renderer.cpp:529-530
    m_draw_objs.push_back(d);
  }
  m_vertices.insert(m_vertices.end(), mesh.vertices.begin(), mesh.vertices.end());
  m_indices.insert(m_indices.end(), mesh.indices.begin(), mesh.indices.end());
  m_geometry_dirty = true;
  m_models.push_back(model);
  return m_models.size() - 1;
}

material_id_t Renderer::add_material(material_data_t& material_data) {
  material_t m;
  std::memset(&m, 0, sizeof m);
  m.ubo = (unsigned)m_materials.size();
  m.num_textures = 0;
  m_materials.push_back(m);
  m_material_data.push_back(material_data);
  m_materials_dirty = true;
  return m_materials.size() - 1;
}

void Renderer::upload_material_data(material_data_t& material_data, material_id_t material_id) {
  if (material_id >= m_material_data.size()) return;
  m_material_data[material_id] = material_data;   // the reference leaks a fresh UBO per call (renderer.cpp:643-651); one array slot here
  m_materials_dirty = true;
}

void Renderer::set_model_material(material_id_t material_id, model_id_t model_id) {
  if (model_id >= m_models.size() || material_id >= m_materials.size()) return;
  const model_t& model = m_models[model_id];
  for (size_t i = 0; i < model.draw_obj_range.size; i++) m_draw_objs[i + model.draw_obj_range.start].material_id = material_id;
}

vec3 Renderer::get_model_dimensions(model_id_t model) { return model < m_models.size() ? m_models[model].dimensions : vec3{0, 0, 0}; }

void Renderer::set_grid_resolution(unsigned int res) {
  m_resolution = res;
  if (!m_device.ok()) return;
  if (m_grid) { vct_grid_destroy(m_grid); m_grid = nullptr; }
  // always 7 levels in the reference (renderer.cpp:186); glTexStorage3D rejects more than log2(res)+1, so clamp.
  // set_grid_storage() can ask for the full chain (m_grid_levels = 0) and for RGBA16F texels instead (BASELINE config 5's variant).
  const int want = m_grid_levels > 0 ? m_grid_levels : (m_grid_format == VCT_GRID_RGBA16F ? VCT_MAX_LEVELS_HOST : 7);
  int levels = 1;
  while (levels < want && (res >> levels) >= 1) levels++;
  check(vct_grid_create_ex(m_device.handle(), (int)res, levels, m_grid_format, &m_grid), "vct_grid_create_ex");
}

void Renderer::set_grid_storage(int vct_grid_format, int levels) {
  m_grid_format = vct_grid_format; m_grid_levels = levels;
  if (m_resolution) set_grid_resolution((unsigned int)m_resolution);
}

void Renderer::set_grid_size(float size) { m_cube_size = size; }

material_t& Renderer::get_material(material_id_t material_id) { return m_materials[material_id]; }

void Renderer::set_rendering_phases(bool direct, bool diffuse, bool specular, bool shadow) {
  m_enable_shadows = shadow; m_enable_direct = direct; m_enable_indirect_diffuse = diffuse; m_enable_indirect_specular = specular;
}

void Renderer::set_voxel_view_dir(int dir, float lod) { m_view_voxel_dir = dir; m_view_voxel_lod = lod; }

void Renderer::queue_model(model_id_t model_id) {
  if (model_id >= m_models.size()) return;     // renderer.cpp:211-212
  m_draw_queue.push_back(m_models[model_id]);  // a COPY: later set_model_transform calls do not affect this frame
}

void Renderer::queue_point_light(point_light_t& point_light) { m_point_lights.push_back(point_light); }

bool Renderer::upload_geometry() {
  if (!m_geometry_dirty) return true;
  m_geometry_dirty = false;
  return check(vct_scene_set_geometry(m_scene, m_vertices.data(), (uint32_t)m_vertices.size(), m_indices.data(), (uint32_t)m_indices.size()),
               "vct_scene_set_geometry");
}

bool Renderer::upload_materials() {
  if (!m_materials_dirty) return true;
  m_materials_dirty = false;
  static_assert(sizeof(material_data_t) == sizeof(vct_material_t), "material_data_t crosses the ABI as vct_material_t");
  return check(vct_scene_set_materials(m_scene, reinterpret_cast<const vct_material_t*>(m_material_data.data()), (uint32_t)m_material_data.size()),
               "vct_scene_set_materials");
}

// draw_models (renderer.cpp:240-257): the queue in order, each model's draw objects in order -> one flat draw list
void Renderer::draw_models() {
  std::vector<vct_draw_t> draws;
  for (const model_t& model : m_draw_queue)
    for (size_t i = 0; i < model.draw_obj_range.size; i++) {
      const draw_obj_t& o = m_draw_objs[model.draw_obj_range.start + i];
      vct_draw_t d;
      d.first_index = (uint32_t)(model.index_base + o.range.start);
      d.index_count = (uint32_t)o.range.size;
      d.vertex_base = (uint32_t)model.vertex_base;
      d.material = (uint32_t)o.material_id;
      std::memcpy(d.model, model.model_matrix.m, sizeof d.model);
      draws.push_back(d);
    }
  check(vct_scene_set_draws(m_scene, draws.data(), (uint32_t)draws.size()), "vct_scene_set_draws");
}

void Renderer::upload_lights() {
  static_assert(sizeof(point_light_t) == sizeof(vct_point_light_t), "point_light_t crosses the ABI as vct_point_light_t");
  check(vct_scene_set_lights(m_scene, reinterpret_cast<const vct_point_light_t*>(m_point_lights.data()), (uint32_t)m_point_lights.size()),
        "vct_scene_set_lights");
}

vct_trace_params_t Renderer::trace_params() const {
  vct_trace_params_t p;
  std::memset(&p, 0, sizeof p);
  p.enable_direct = m_enable_direct; p.enable_diffuse = m_enable_indirect_diffuse;
  p.enable_specular = m_enable_indirect_specular; p.enable_shadow = m_enable_shadows;
  p.view_voxel_dir = m_view_voxel_dir; p.view_voxel_lod = m_view_voxel_lod;
  p.n_diffuse_cones = m_diffuse_cones;
  p.tile_rank = m_rank; p.tile_nranks = m_nranks;
  p.sampler = m_sampler;
  return p;
}

// voxelize() (renderer.cpp:316-353): clear the voxel textures, scatter the scene into them, build the mip chains
bool Renderer::voxelize() {
  const int R = (int)m_resolution;
  if ((m_last_rc = vct_grid_clear(m_grid)) != VCT_OK) return false;
  if ((m_last_rc = vct_voxelize(m_device.handle(), m_scene, m_grid, 0, R)) != VCT_OK) return false;
  return filter();
}
bool Renderer::filter() { return (m_last_rc = vct_mipmap(m_device.handle(), m_grid)) == VCT_OK; }
// visualize() (renderer.cpp:355-390): visibility pass + the fragment shader's cone tracing
bool Renderer::visualize() {
  const vct_trace_params_t p = trace_params();
  if ((m_last_rc = vct_gbuffer(m_device.handle(), m_scene, m_camera.view.m, m_camera.projection.m, m_target)) != VCT_OK) return false;
  return (m_last_rc = vct_cone_trace(m_device.handle(), m_scene, m_grid, m_camera.view.m, &p, m_target)) == VCT_OK;
}

void Renderer::render() {
  if (!ok()) {
    std::fprintf(stderr, "vct::Renderer::render: no usable sm_100 device (%s); there is no CPU fallback\n", m_device.error());
    m_draw_queue.clear(); m_point_lights.clear();
    return;
  }
  upload_geometry();
  upload_materials();
  check(vct_scene_set_cube_size(m_scene, m_cube_size), "vct_scene_set_cube_size");
  draw_models();
  upload_lights();
  const vct_trace_params_t p = trace_params();
  // clear -> voxelize -> filter -> visualize (renderer.cpp:392-402), asynchronous on the device stream.  VCT_ERR_OVERFLOW = an
  // earlier frame's voxelization did not fit the fragment arena; the library has grown it, the frame is simply issued again.
  for (int attempt = 0; attempt < 4; attempt++) {
    if (m_staged) { if (voxelize()) visualize(); }
    else m_last_rc = vct_render_frame(m_device.handle(), m_scene, m_grid, m_target, m_camera.view.m, m_camera.projection.m, &p);
    if (m_last_rc != VCT_ERR_OVERFLOW) break;
  }
  check(m_last_rc, m_staged ? "voxelize/visualize" : "vct_render_frame");
  m_draw_queue.clear();      // renderer.cpp:403-404
  m_point_lights.clear();
}

bool Renderer::read_frame(uint32_t* rgba8) { return ok() && check(vct_target_download_frame(m_target, rgba8), "vct_target_download_frame"); }
uint64_t Renderer::read_frame_async(uint32_t* pinned_rgba8) {
  uint64_t ticket = 0;
  if (!ok() || !check(vct_target_download_frame_async(m_target, pinned_rgba8, &ticket), "vct_target_download_frame_async")) return 0;
  return ticket;
}
bool Renderer::wait_frame(uint64_t ticket) { return ok() && check(vct_target_download_wait(m_target, ticket), "vct_target_download_wait"); }
void* Renderer::frame_device_ptr() { return m_target ? vct_target_frame_device_ptr(m_target) : nullptr; }
bool Renderer::read_voxels(int level, int dir, uint32_t* rgba8) { return ok() && check(vct_grid_download(m_grid, level, dir, rgba8), "vct_grid_download"); }

}  // namespace vct
