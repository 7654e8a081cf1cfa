// obj_loader.cpp -- OBJ/MTL reader behind Renderer::load_model.
//
// The reference parses with the vendored tinyobjloader v1.1.0 (triangulate = true) and then flattens /
// dedupes / splits per material in Renderer::load_model (src/renderer.cpp:407-557).  This is an own reader
// that yields the same streams for the subset of OBJ/MTL the reference's assets and tinyobj's material
// model cover (tests pin it against tinyobjloader compiled from the reference tree, oracle/_ref):
//   * v / vn / vt, f with v, v/vt, v//vn, v/vt/vn corners, negative (relative) indices, polygons
//     fan-triangulated (corner 0, k, k+1) like tinyobj's triangulate path
//   * a new range ("draw object") starts at every g / o statement and wherever the material changes inside
//     a group (renderer.cpp:519-556); usemtl alone does not start a new shape (tiny_obj_loader.h usemtl case)
//   * MTL: InitMaterial defaults (tiny_obj_loader.h:936-957: everything 0 except dissolve = 1, shininess = 1,
//     ior = 1), Ka Kd Ks Kt/Tf Ke Ni Ns illum d Tr (d wins over Tr, :1203-1222) Pr Pm Ps Pc Pcr aniso anisor
//   * one vertex per distinct (position, normal, texcoord) triple in first-use order (renderer.cpp:460-515)
#include <cerrno>
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <unordered_map>

#include "vct/renderer.h"

namespace vct {
namespace {

struct VertexKey {
  vct_vertex_t v;
  bool operator==(const VertexKey& o) const { return std::memcmp(&v, &o.v, sizeof v) == 0; }
};
struct VertexKeyHash {
  size_t operator()(const VertexKey& k) const {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&k.v);
    uint64_t h = 1469598103934665603ull;   // FNV-1a over the 8 words
    for (int i = 0; i < 8; i++) { h ^= w[i]; h *= 1099511628211ull; }
    return (size_t)h;
  }
};

std::vector<std::string> split_ws(const std::string& line) {
  std::vector<std::string> tok;
  std::istringstream is(line);
  std::string t;
  while (is >> t) tok.push_back(t);
  return tok;
}

float to_f(const std::string& s) { return (float)std::strtod(s.c_str(), nullptr); }

float arg_f(const std::vector<std::string>& tok, size_t i) { return i < tok.size() ? to_f(tok[i]) : 0.0f; }

material_data_t default_material() {
  material_data_t m;
  std::memset(&m, 0, sizeof m);
  m.dissolve = 1.0f; m.shininess = 1.0f; m.ior = 1.0f;
  return m;
}

bool parse_mtl(const std::string& path, std::vector<material_data_t>* mats, std::vector<std::string>* names) {
  std::ifstream f(path);
  if (!f) return false;
  std::string raw;
  bool have = false, has_d = false;
  material_data_t cur = default_material();
  while (std::getline(f, raw)) {
    const size_t hash = raw.find('#');
    if (hash != std::string::npos) raw.erase(hash);
    const std::vector<std::string> tok = split_ws(raw);
    if (tok.empty()) continue;
    const std::string& key = tok[0];
    if (key == "newmtl") {
      if (have) mats->push_back(cur);
      cur = default_material();
      names->push_back(tok.size() > 1 ? tok[1] : "");
      have = true; has_d = false;
      continue;
    }
    if (!have) continue;
    auto set3 = [&](vec4& d) { d.x = arg_f(tok, 1); d.y = arg_f(tok, 2); d.z = arg_f(tok, 3); };
    if (key == "Ka") set3(cur.ambient);
    else if (key == "Kd") set3(cur.diffuse);
    else if (key == "Ks") set3(cur.specular);
    else if (key == "Kt" || key == "Tf") set3(cur.transmittance);
    else if (key == "Ke") { cur.emission.x = arg_f(tok, 1); cur.emission.y = arg_f(tok, 2); cur.emission.z = arg_f(tok, 3); }
    else if (key == "Ni") cur.ior = arg_f(tok, 1);
    else if (key == "Ns") cur.shininess = arg_f(tok, 1);
    else if (key == "illum") cur.illum = (int)arg_f(tok, 1);
    else if (key == "d") { cur.dissolve = arg_f(tok, 1); has_d = true; }
    else if (key == "Tr") { if (!has_d) cur.dissolve = 1.0f - arg_f(tok, 1); }
    else if (key == "Pr") cur.roughness = arg_f(tok, 1);
    else if (key == "Pm") cur.metallic = arg_f(tok, 1);
    else if (key == "Ps") cur.sheen = arg_f(tok, 1);
    else if (key == "Pc") cur.clearcoat_thickness = arg_f(tok, 1);
    else if (key == "Pcr") cur.clearcoat_roughness = arg_f(tok, 1);
    else if (key == "aniso") cur.anisotropy = arg_f(tok, 1);
    else if (key == "anisor") cur.anisotropy_rotation = arg_f(tok, 1);
  }
  if (have) mats->push_back(cur);
  return true;
}

struct Corner { int v, t, n; };

// "v", "v/t", "v//n", "v/t/n"; 1-based, negative = relative to the end
bool parse_corner(const std::string& s, size_t nv, size_t nt, size_t nn, Corner* c) {
  int idx[3] = {0, 0, 0};
  bool present[3] = {false, false, false};
  size_t start = 0;
  for (int k = 0; k < 3 && start <= s.size(); k++) {
    size_t end = s.find('/', start);
    if (end == std::string::npos) end = s.size();
    if (end > start) { idx[k] = std::atoi(s.substr(start, end - start).c_str()); present[k] = true; }
    start = end + 1;
  }
  if (!present[0]) return false;
  auto fix = [](int i, size_t n) { return i > 0 ? i - 1 : (int)n + i; };
  c->v = fix(idx[0], nv);
  c->t = present[1] ? fix(idx[1], nt) : -1;
  c->n = present[2] ? fix(idx[2], nn) : -1;
  return c->v >= 0 && (size_t)c->v < nv && (c->t < 0 || (size_t)c->t < nt) && (c->n < 0 || (size_t)c->n < nn);
}

}  // namespace

// VCTMESH1: the repo's binary mesh fixture (voxel_cone_tracing_b200/scene.py save_vctmesh) = the streams this
// reader produces for an OBJ, stored so that the benchmark scenes exist where the OBJ files do not
static bool load_vctmesh(const char* filename, loaded_mesh_t* out) {
  FILE* f = std::fopen(filename, "rb");
  if (!f) return false;
  char magic[8];
  uint32_t hdr[4];
  bool ok = std::fread(magic, 1, 8, f) == 8 && std::memcmp(magic, "VCTMESH1", 8) == 0 && std::fread(hdr, 4, 4, f) == 4;
  if (ok) {
    out->vertices.resize(hdr[0]); out->indices.resize(hdr[1]); out->ranges.resize(hdr[2]); out->materials.resize(hdr[3]);
    ok = std::fread(out->vertices.data(), sizeof(vct_vertex_t), hdr[0], f) == hdr[0] && std::fread(out->indices.data(), 4, hdr[1], f) == hdr[1];
    for (uint32_t i = 0; ok && i < hdr[2]; i++) {
      uint32_t r[3];
      ok = std::fread(r, 4, 3, f) == 3;
      out->ranges[i] = {r[0], r[1], (int)r[2]};
    }
    ok = ok && std::fread(out->materials.data(), sizeof(material_data_t), hdr[3], f) == hdr[3];
  }
  std::fclose(f);
  if (!ok) return false;
  out->bbox_min = {FLT_MAX, FLT_MAX, FLT_MAX};
  out->bbox_max = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (uint32_t i : out->indices) {
    if (i >= out->vertices.size()) return false;
    const float* p = out->vertices[i].pos;
    out->bbox_min = {std::fmin(out->bbox_min.x, p[0]), std::fmin(out->bbox_min.y, p[1]), std::fmin(out->bbox_min.z, p[2])};
    out->bbox_max = {std::fmax(out->bbox_max.x, p[0]), std::fmax(out->bbox_max.y, p[1]), std::fmax(out->bbox_max.z, p[2])};
  }
  return !out->indices.empty();
}

bool load_obj(const char* filename, loaded_mesh_t* out) {
  const size_t len = std::strlen(filename);
  if (len > 8 && std::strcmp(filename + len - 8, ".vctmesh") == 0) return load_vctmesh(filename, out);
  std::ifstream f(filename);
  if (!f) return false;
  const std::string path = filename;
  const std::string base = path.substr(0, path.find_last_of("\\/") + 1);   // renderer.cpp:413-414
  std::vector<vec3> pos, nrm;
  std::vector<vec2> tex;
  std::vector<std::string> names;
  std::unordered_map<VertexKey, uint32_t, VertexKeyHash> uniq;
  out->vertices.clear(); out->indices.clear(); out->ranges.clear(); out->materials.clear();
  out->bbox_min = {FLT_MAX, FLT_MAX, FLT_MAX};
  out->bbox_max = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  int cur_mat = -1, range_mat = -1;
  bool range_open = false;
  size_t start = 0;
  auto close_range = [&]() {
    if (out->indices.size() > start) out->ranges.push_back({(uint32_t)start, (uint32_t)(out->indices.size() - start), range_open ? range_mat : -1});
    start = out->indices.size();
    range_open = false;
  };
  std::string raw;
  std::vector<Corner> corners;
  while (std::getline(f, raw)) {
    const size_t hash = raw.find('#');
    if (hash != std::string::npos) raw.erase(hash);
    const std::vector<std::string> tok = split_ws(raw);
    if (tok.empty()) continue;
    const std::string& key = tok[0];
    if (key == "v") pos.push_back({arg_f(tok, 1), arg_f(tok, 2), arg_f(tok, 3)});
    else if (key == "vn") nrm.push_back({arg_f(tok, 1), arg_f(tok, 2), arg_f(tok, 3)});
    else if (key == "vt") tex.push_back({arg_f(tok, 1), arg_f(tok, 2)});
    else if (key == "mtllib") {
      if (tok.size() > 1 && !parse_mtl(base + tok[1], &out->materials, &names))
        std::fprintf(stderr, "load_obj: material file %s not found\n", (base + tok[1]).c_str());   // tinyobj warns and continues
    } else if (key == "usemtl") {
      cur_mat = -1;
      for (size_t i = 0; i < names.size(); i++) if (tok.size() > 1 && names[i] == tok[1]) { cur_mat = (int)i; break; }
    } else if (key == "g" || key == "o") {
      close_range();
    } else if (key == "f") {
      corners.clear();
      bool ok = true;
      for (size_t i = 1; i < tok.size(); i++) {
        Corner c;
        if (!parse_corner(tok[i], pos.size(), tex.size(), nrm.size(), &c)) { ok = false; break; }
        corners.push_back(c);
      }
      if (!ok || corners.size() < 3) continue;
      if (range_open && range_mat != cur_mat) close_range();
      range_mat = cur_mat; range_open = true;
      for (size_t k = 1; k + 1 < corners.size(); k++) {
        const Corner tri[3] = {corners[0], corners[k], corners[k + 1]};
        for (const Corner& c : tri) {
          VertexKey key_v;
          std::memset(&key_v, 0, sizeof key_v);
          const vec3 p = pos[c.v];
          key_v.v.pos[0] = p.x; key_v.v.pos[1] = p.y; key_v.v.pos[2] = p.z;
          if (c.n >= 0) { key_v.v.norm[0] = nrm[c.n].x; key_v.v.norm[1] = nrm[c.n].y; key_v.v.norm[2] = nrm[c.n].z; }
          if (c.t >= 0) { key_v.v.uv[0] = tex[c.t].x; key_v.v.uv[1] = tex[c.t].y; }
          const vct_vertex_t as_read = key_v.v;
          float* kf = reinterpret_cast<float*>(&key_v.v);
          for (int a = 0; a < 8; a++)   // -0.0f and 0.0f compare equal in the reference's operator== (renderer.cpp:31-34): one key, first-seen bits kept
            if (kf[a] == 0.0f) kf[a] = 0.0f;
          out->bbox_min = {std::fmin(out->bbox_min.x, p.x), std::fmin(out->bbox_min.y, p.y), std::fmin(out->bbox_min.z, p.z)};
          out->bbox_max = {std::fmax(out->bbox_max.x, p.x), std::fmax(out->bbox_max.y, p.y), std::fmax(out->bbox_max.z, p.z)};
          auto it = uniq.find(key_v);
          uint32_t j;
          if (it == uniq.end()) {
            j = (uint32_t)out->vertices.size();
            uniq.emplace(key_v, j);
            out->vertices.push_back(as_read);
          } else j = it->second;
          out->indices.push_back(j);
        }
      }
    }
  }
  close_range();
  return !out->indices.empty();
}

}  // namespace vct
