// vct_demo.cpp -- the reference's main() scene (src/main.cpp:82-121,369-384) without the window: builds the
// Cornell box (+ optionally the refractive Suzanne) through the Renderer API, renders N frames, prints the
// per-stage device timings and writes the last frame.  Also serves the tests: --dump writes every array that
// crossed the C ABI so that the Python harness / oracle can re-render exactly the same inputs.
//
//   vct_demo [--assets DIR] [--res 128] [--size 800x600] [--suzanne] [--theta 0.0] [--frames 1] [--sampler 1]
//            [--cones 9] [--ppm out.ppm] [--raw out.rgba] [--dump scene.bin] [--view-dir D --view-lod L]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "vct/renderer.h"

using namespace vct;

// the OBJ when it is there (a checkout of the reference's assets), else the committed binary fixture of the same mesh
static std::string asset(const std::string& dir, const char* obj, const char* fixture) {
  const std::string a = dir + "/" + obj;
  FILE* f = std::fopen(a.c_str(), "rb");
  if (f) { std::fclose(f); return a; }
  return dir + "/" + fixture;
}

static void write_ppm(const char* path, const std::vector<uint32_t>& px, int W, int H) {
  FILE* f = std::fopen(path, "wb");
  if (!f) return;
  std::fprintf(f, "P6\n%d %d\n255\n", W, H);
  for (int j = H - 1; j >= 0; j--)   // frame rows are bottom-up (GL window coordinates)
    for (int i = 0; i < W; i++) {
      const uint32_t p = px[(size_t)j * W + i];
      const unsigned char rgb[3] = {(unsigned char)(p & 255), (unsigned char)((p >> 8) & 255), (unsigned char)((p >> 16) & 255)};
      std::fwrite(rgb, 1, 3, f);
    }
  std::fclose(f);
}

int main(int argc, char** argv) {
  std::string assets = "assets", ppm, raw, dump;
  int res = 128, W = 800, H = 600, frames = 1, sampler = VCT_SAMPLER_TEX, cones = 9, view_dir = 7;
  float theta = 0.0f, view_lod = 0.0f;
  bool suzanne = false;
  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    auto next = [&]() { return i + 1 < argc ? argv[++i] : ""; };
    if (a == "--assets") assets = next();
    else if (a == "--res") res = std::atoi(next());
    else if (a == "--size") { if (std::sscanf(next(), "%dx%d", &W, &H) != 2) { std::fprintf(stderr, "bad --size\n"); return 2; } }
    else if (a == "--suzanne") suzanne = true;
    else if (a == "--theta") theta = (float)std::atof(next());
    else if (a == "--frames") frames = std::atoi(next());
    else if (a == "--sampler") sampler = std::atoi(next());
    else if (a == "--cones") cones = std::atoi(next());
    else if (a == "--ppm") ppm = next();
    else if (a == "--raw") raw = next();
    else if (a == "--dump") dump = next();
    else if (a == "--view-dir") view_dir = std::atoi(next());
    else if (a == "--view-lod") view_lod = (float)std::atof(next());
    else { std::fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
  }

  Renderer renderer(W, H);
  if (!renderer.ok()) { std::fprintf(stderr, "vct_demo: no usable B200 (%s)\n", renderer.device().error()); return 3; }
  renderer.set_sampler(sampler);
  renderer.set_diffuse_cone_count(cones);
  renderer.set_grid_resolution((unsigned)res);
  renderer.set_voxel_view_dir(view_dir, view_lod);

  // Create scene (main.cpp:85-103)
  model_id_t box = renderer.load_model(asset(assets, "CornellBox-Glossy.obj", "cornell_glossy.vctmesh").c_str());
  if (box == INVALID_ID) return 4;
  model_id_t dynamic_object = INVALID_ID;
  if (suzanne) {
    dynamic_object = renderer.load_model(asset(assets, "suzanne.obj", "suzanne.vctmesh").c_str());
    if (dynamic_object == INVALID_ID) return 4;
    material_data_t m;
    std::memset(&m, 0, sizeof m);   // the reference leaves the PBR fields uninitialised; no shader reads them
    m.ambient = {1, 1, 1, 1};
    m.diffuse = {0, 0, 0, 0};
    m.specular = {1, 1, 1, 1};
    m.transmittance = {1, 1, 1, 1};
    m.emission = {0.0f, 0.0f, 0.25f};
    m.shininess = 1000; m.ior = 5; m.dissolve = 0.1f; m.illum = 4;
    material_id_t id = renderer.add_material(m);
    renderer.set_model_material(id, dynamic_object);
  }
  renderer.set_grid_size(3);   // main.cpp:105

  // Camera camera(glm::vec3(0, .9, 3), 0, -90); set_perspective(45.0f /* radians in glm 0.9.9 */, W/H, 0.1, 100) (main.cpp:107-109)
  const vec3 eye = {0.0f, 0.9f, 3.0f};
  const vec3 front = camera_front(0.0f, -90.0f);
  mat4 view = look_at(eye, eye + front, {0, 1, 0});
  mat4 proj = perspective(45.0f, (float)W / (float)H, 0.1f, 100.0f);
  renderer.set_camera_transform(view, proj);

  point_light_t ceiling_light;   // main.cpp:115-118
  ceiling_light.position = {0.0f, 1.4f, 0.0f};
  ceiling_light.color = {1.0f, 1.0f, 1.0f};
  ceiling_light.intensity = 1.0f;

  float acc[8] = {0};
  int timed = 0;
  mat4 dynamic_object_matrix = mat4::identity();
  for (int f = 0; f < frames; f++) {
    if (suzanne) {   // main.cpp:369-374
      dynamic_object_matrix = mat4::identity();
      dynamic_object_matrix = translate(dynamic_object_matrix, {0.0f, 1.1f, -0.5f});
      dynamic_object_matrix = rotate(dynamic_object_matrix, theta + 0.05f * (float)f, {0, 1, 0});
      dynamic_object_matrix = scale(dynamic_object_matrix, {0.3f, 0.3f, 0.3f});
      renderer.set_model_transform(dynamic_object, dynamic_object_matrix);
    }
    renderer.set_model_transform(box, mat4::identity());
    renderer.queue_model(box);
    if (suzanne) renderer.queue_model(dynamic_object);
    renderer.queue_point_light(ceiling_light);
    renderer.render();
    float t[8];
    if (f >= frames / 2 && renderer.device().last_frame_timings(t)) {   // second half: warmed up
      for (int k = 0; k < 8; k++) acc[k] += t[k];
      timed++;
    }
  }
  std::vector<uint32_t> frame((size_t)W * H);
  if (!renderer.read_frame(frame.data())) return 5;
  if (timed)
    std::printf("frames=%d res=%d size=%dx%d triangles=%zu  clear=%.1fus voxelize=%.1fus mipmap=%.1fus gbuffer=%.1fus trace=%.1fus total=%.1fus\n", frames, res,
                W, H, renderer.triangle_count(), 1e3 * acc[0] / timed, 1e3 * acc[1] / timed, 1e3 * acc[2] / timed, 1e3 * acc[3] / timed,
                1e3 * acc[4] / timed, 1e3 * acc[5] / timed);
  if (!ppm.empty()) write_ppm(ppm.c_str(), frame, W, H);
  if (!raw.empty()) {
    FILE* fo = std::fopen(raw.c_str(), "wb");
    if (fo) { std::fwrite(frame.data(), 4, frame.size(), fo); std::fclose(fo); }
  }
  if (!dump.empty()) {
    // camera + last dynamic-object matrix so that the harness can rebuild the exact inputs of the last frame
    FILE* fo = std::fopen(dump.c_str(), "wb");
    if (fo) {
      std::fwrite(view.m, 4, 16, fo);
      std::fwrite(proj.m, 4, 16, fo);
      std::fwrite(dynamic_object_matrix.m, 4, 16, fo);
      std::fclose(fo);
    }
  }
  return 0;
}
