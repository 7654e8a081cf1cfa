// obj_dump.cpp -- runs vct::load_obj (the reader behind Renderer::load_model) on one OBJ and writes the
// result as a VCTMESH1 file (the binary fixture format of voxel_cone_tracing_b200/scene.py).  CPU-only;
// tests compare it with the Python reader and with tinyobjloader's streams.
#include <cstdio>
#include <cstring>

#include "vct/renderer.h"

int main(int argc, char** argv) {
  if (argc != 3) { std::fprintf(stderr, "usage: obj_dump in.obj out.vctmesh\n"); return 2; }
  vct::loaded_mesh_t m;
  if (!vct::load_obj(argv[1], &m)) { std::fprintf(stderr, "Error loading file %s\n", argv[1]); return 1; }
  FILE* f = std::fopen(argv[2], "wb");
  if (!f) return 1;
  const uint32_t hdr[4] = {(uint32_t)m.vertices.size(), (uint32_t)m.indices.size(), (uint32_t)m.ranges.size(), (uint32_t)m.materials.size()};
  std::fwrite("VCTMESH1", 1, 8, f);
  std::fwrite(hdr, 4, 4, f);
  std::fwrite(m.vertices.data(), sizeof(vct_vertex_t), m.vertices.size(), f);
  std::fwrite(m.indices.data(), 4, m.indices.size(), f);
  for (const auto& r : m.ranges) {
    std::fwrite(&r.first_index, 4, 1, f); std::fwrite(&r.index_count, 4, 1, f); std::fwrite(&r.material, 4, 1, f);
  }
  std::fwrite(m.materials.data(), sizeof(vct::material_data_t), m.materials.size(), f);
  std::fclose(f);
  std::printf("%zu vertices, %zu indices, %zu ranges, %zu materials, bbox (%g %g %g)-(%g %g %g)\n", m.vertices.size(), m.indices.size(), m.ranges.size(),
              m.materials.size(), m.bbox_min.x, m.bbox_min.y, m.bbox_min.z, m.bbox_max.x, m.bbox_max.y, m.bbox_max.z);
  return 0;
}
