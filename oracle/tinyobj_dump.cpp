// tinyobj_dump.cpp -- TEST INFRASTRUCTURE.  Loads an OBJ with the reference's vendored tinyobjloader
// (compiled from /root/reference/thirdparty/tinyobjloader where it lies, never copied) exactly the way
// Renderer::load_model does (src/renderer.cpp:417: triangulate = true) and dumps the flattened
// per-index vertex stream, the per-face material ids and the material constants that
// create_material (src/renderer.cpp:49-81) forwards to the shaders.  tests/test_scene_inputs.py
// compares this with our own OBJ/MTL reader and with the committed assets/*.vctmesh fixtures.
//
// Output (text, one record per line):
//   shapes <n> materials <m>
//   M <i> Ka3 Kd3 Ks3 Tf3 Ke3 Ns Ni d illum
//   S <shape index> <n_indices>
//   V px py pz nx ny nz u v material_id      (one per index, floats printed as hex bit patterns)
#define TINYOBJLOADER_IMPLEMENTATION
#include <tiny_obj_loader.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

static unsigned bits(float f) { unsigned u; memcpy(&u, &f, 4); return u; }

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: tinyobj_dump file.obj\n"); return 2; }
  std::string file = argv[1];
  std::string base = file.substr(0, file.find_last_of("\\/") + 1);
  tinyobj::attrib_t attrib;
  std::vector<tinyobj::shape_t> shapes;
  std::vector<tinyobj::material_t> materials;
  std::string err;
  if (!tinyobj::LoadObj(&attrib, &shapes, &materials, &err, file.c_str(), base.c_str(), true)) {
    fprintf(stderr, "load failed: %s\n", err.c_str());
    return 1;
  }
  printf("shapes %zu materials %zu\n", shapes.size(), materials.size());
  for (size_t i = 0; i < materials.size(); i++) {
    const tinyobj::material_t& m = materials[i];
    printf("M %zu", i);
    for (int k = 0; k < 3; k++) printf(" %08x", bits(m.ambient[k]));
    for (int k = 0; k < 3; k++) printf(" %08x", bits(m.diffuse[k]));
    for (int k = 0; k < 3; k++) printf(" %08x", bits(m.specular[k]));
    for (int k = 0; k < 3; k++) printf(" %08x", bits(m.transmittance[k]));
    for (int k = 0; k < 3; k++) printf(" %08x", bits(m.emission[k]));
    printf(" %08x %08x %08x %d\n", bits(m.shininess), bits(m.ior), bits(m.dissolve), m.illum);
  }
  for (size_t s = 0; s < shapes.size(); s++) {
    const tinyobj::mesh_t& mesh = shapes[s].mesh;
    printf("S %zu %zu\n", s, mesh.indices.size());
    for (size_t i = 0; i < mesh.indices.size(); i++) {
      const tinyobj::index_t& ix = mesh.indices[i];
      float p[3] = {0, 0, 0}, n[3] = {0, 0, 0}, t[2] = {0, 0};
      for (int k = 0; k < 3; k++) p[k] = attrib.vertices[3 * ix.vertex_index + k];
      if (ix.normal_index >= 0) for (int k = 0; k < 3; k++) n[k] = attrib.normals[3 * ix.normal_index + k];
      if (ix.texcoord_index >= 0) for (int k = 0; k < 2; k++) t[k] = attrib.texcoords[2 * ix.texcoord_index + k];
      printf("V %08x %08x %08x %08x %08x %08x %08x %08x %d\n", bits(p[0]), bits(p[1]), bits(p[2]), bits(n[0]), bits(n[1]), bits(n[2]),
             bits(t[0]), bits(t[1]), mesh.material_ids[i / 3]);
    }
  }
  return 0;
}
