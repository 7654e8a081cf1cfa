// obj_loader.cpp -- OBJ/MTL reader behind Renderer::load_model.
//
// The reference parses with the vendored tinyobjloader v1.1.0 (triangulate = true) and then flattens /
// dedupes / splits per material in Renderer::load_model (src/renderer.cpp:407-557).  This is an own reader
// that restates the loader's grammar so that a model is what the reference would have loaded, also for files
// an exporter would not write (tests/test_obj_reader_fuzz.py holds it against tinyobjloader compiled from the
// reference tree, oracle/_ref, on random files):
//   * numbers: the loader's own float parser (tiny_obj_loader.h:498-605: [sign] digits [. digits] [e [sign] digits],
//     value of the conforming prefix, 0 where it does not parse; decimal digits added through a table of powers,
//     10^e as ldexp(m * 5^e, e)) -- not strtod
//   * v / vn / vt, f with v, v/vt, v//vn, v/vt/vn corners read the way parseTriple does (:745-799), negative
//     (relative) indices, a zero index fails the load, polygons fan-triangulated (corner 0, k-1, k)
//   * faces collect in a pending group; `usemtl` with a different material moves the group into the current shape,
//     `g` / `o` do the same and close the shape: `g` keeps it if it has triangles, `o` only if the pending group was
//     not empty (so `usemtl` directly in front of an `o` loses the shape, as in the loader); a material name is the
//     rest of the line
//   * a new range ("draw object") per shape and per run of one material inside it (renderer.cpp:519-556)
//   * MTL (LoadMtl, :1049-1431): InitMaterial defaults (everything 0 except dissolve = 1, shininess = 1, ior = 1),
//     Ka Kd Ks Kt/Tf Ke Ni Ns illum d Tr (d wins over Tr) Pr Pm Ps Pc Pcr aniso anisor; statements in front of the
//     first newmtl are dropped with their nameless material, a file without newmtl yields that one nameless material
//   * one vertex per distinct (position, normal, texcoord) triple in first-use order (renderer.cpp:460-515)
#include <cerrno>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
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

inline bool is_blank(char c) { return c == ' ' || c == '\t'; }
inline bool is_digit(char c) { return c >= '0' && c <= '9'; }

// \n, \r\n and a lone \r all end a line (the loader's safeGetline)
std::vector<std::string> read_lines(std::istream& f) {
  std::stringstream ss;
  ss << f.rdbuf();
  const std::string blob = ss.str();
  std::vector<std::string> lines;
  std::string cur;
  for (size_t i = 0; i < blob.size(); i++) {
    const char c = blob[i];
    if (c == '\r' || c == '\n') {
      if (c == '\r' && i + 1 < blob.size() && blob[i + 1] == '\n') i++;
      lines.push_back(cur);
      cur.clear();
    } else cur.push_back(c);
  }
  lines.push_back(cur);
  return lines;
}

// value of the longest conforming prefix of [s, e); false where the loader reports a parse failure
bool try_parse_double(const char* s, const char* e, double* out) {
  if (s >= e) return false;
  static const double lut[8] = {1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001};
  double sign = 1.0;
  if (*s == '+' || *s == '-') { sign = *s == '-' ? -1.0 : 1.0; s++; }
  else if (!is_digit(*s)) return false;
  double mant = 0.0;
  int read = 0;
  while (s < e && is_digit(*s)) { mant = mant * 10.0 + (*s - '0'); s++; read++; }
  if (read == 0) return false;
  int exponent = 0;
  if (s < e && *s == '.') {
    s++;
    read = 1;
    while (s < e && is_digit(*s)) {
      mant += (*s - '0') * (read < 8 ? lut[read] : std::pow(10.0, -read));
      read++; s++;
    }
  }
  if (s < e && (*s == 'e' || *s == 'E')) {
    s++;
    int esign = 1;
    if (s < e && (*s == '+' || *s == '-')) { esign = *s == '-' ? -1 : 1; s++; }
    else if (!(s < e && is_digit(*s))) return false;
    read = 0;
    while (s < e && is_digit(*s)) { exponent = exponent * 10 + (*s - '0'); s++; read++; }
    exponent *= esign;
    if (read == 0) return false;
  }
  *out = sign * (exponent ? std::ldexp(mant * std::pow(5.0, exponent), exponent) : mant);
  return true;
}

size_t skip_blanks(const std::string& l, size_t p) { while (p < l.size() && is_blank(l[p])) p++; return p; }

// next blank-separated token of `l` from *p as a float (0 where it does not parse); *p moves behind the token
float parse_real(const std::string& l, size_t* p) {
  size_t b = skip_blanks(l, *p), e = b;
  while (e < l.size() && !is_blank(l[e]) && l[e] != '\r') e++;
  double v = 0.0;
  try_parse_double(l.c_str() + b, l.c_str() + e, &v);
  *p = e;
  return (float)v;
}

bool is_stmt(const std::string& t, const char* key) {
  const size_t n = std::strlen(key);
  return t.size() > n && t.compare(0, n, key) == 0 && is_blank(t[n]);
}

material_data_t default_material() {
  material_data_t m;
  std::memset(&m, 0, sizeof m);
  m.dissolve = 1.0f; m.shininess = 1.0f; m.ior = 1.0f;
  return m;
}

bool parse_mtl(const std::string& path, std::vector<material_data_t>* mats, std::map<std::string, int>* by_name) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  material_data_t cur = default_material();
  std::string name;
  bool has_d = false;
  auto flush = [&]() {
    by_name->insert(std::make_pair(name, (int)mats->size()));   // the first material of a name keeps it
    mats->push_back(cur);
  };
  for (std::string raw : read_lines(f)) {
    raw.erase(raw.find_last_not_of(" \t") + 1);
    const std::string tok = raw.substr(skip_blanks(raw, 0));
    if (tok.empty() || tok[0] == '#') continue;
    if (is_stmt(tok, "newmtl")) {
      if (!name.empty()) flush();
      cur = default_material();
      name = tok.substr(7);
      has_d = false;
      continue;
    }
    size_t p = 2;
    auto set3 = [&](float* d) { d[0] = parse_real(tok, &p); d[1] = parse_real(tok, &p); d[2] = parse_real(tok, &p); };
    auto one = [&](size_t from) { size_t q = from; return parse_real(tok, &q); };
    if (is_stmt(tok, "Ka")) set3(&cur.ambient.x);
    else if (is_stmt(tok, "Kd")) set3(&cur.diffuse.x);
    else if (is_stmt(tok, "Ks")) set3(&cur.specular.x);
    else if (is_stmt(tok, "Kt") || is_stmt(tok, "Tf")) set3(&cur.transmittance.x);
    else if (is_stmt(tok, "Ke")) set3(&cur.emission.x);
    else if (is_stmt(tok, "Ni")) cur.ior = one(2);
    else if (is_stmt(tok, "Ns")) cur.shininess = one(2);
    else if (is_stmt(tok, "illum")) cur.illum = std::atoi(tok.c_str() + 6);
    else if (is_stmt(tok, "d")) { cur.dissolve = one(1); has_d = true; }
    else if (is_stmt(tok, "Tr")) { if (!has_d) cur.dissolve = 1.0f - one(2); }
    else if (is_stmt(tok, "Pr")) cur.roughness = one(2);
    else if (is_stmt(tok, "Pm")) cur.metallic = one(2);
    else if (is_stmt(tok, "Ps")) cur.sheen = one(2);
    else if (is_stmt(tok, "Pc")) cur.clearcoat_thickness = one(2);
    else if (is_stmt(tok, "Pcr")) cur.clearcoat_roughness = one(3);
    else if (is_stmt(tok, "aniso")) cur.anisotropy = one(5);
    else if (is_stmt(tok, "anisor")) cur.anisotropy_rotation = one(6);
  }
  flush();   // the last material is kept whatever its name
  return true;
}

struct Corner { int v, t, n; };
struct Tri { Corner c[3]; int material; };

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
  std::ifstream f(filename, std::ios::binary);
  if (!f) return false;
  const std::string path = filename;
  const std::string base = path.substr(0, path.find_last_of("\\/") + 1);   // renderer.cpp:413-414
  std::vector<vec3> pos, nrm;
  std::vector<vec2> tex;
  std::map<std::string, int> by_name;
  out->vertices.clear(); out->indices.clear(); out->ranges.clear(); out->materials.clear();
  out->bbox_min = {FLT_MAX, FLT_MAX, FLT_MAX};
  out->bbox_max = {-FLT_MAX, -FLT_MAX, -FLT_MAX};

  // ---- the loader's state machine: pending face group -> current shape -> shapes
  int material = -1;
  std::vector<std::vector<Corner>> group;
  std::vector<Tri> shape;
  std::vector<std::vector<Tri>> shapes;
  auto export_group = [&]() {
    if (group.empty()) return false;
    for (const std::vector<Corner>& face : group)
      for (size_t k = 2; k < face.size(); k++) shape.push_back({{face[0], face[k - 1], face[k]}, material});
    return true;
  };
  for (const std::string& raw : read_lines(f)) {
    const std::string tok = raw.substr(skip_blanks(raw, 0));
    if (tok.empty() || tok[0] == '#') continue;
    size_t p;
    if (is_stmt(tok, "v")) { p = 2; vec3 v; v.x = parse_real(tok, &p); v.y = parse_real(tok, &p); v.z = parse_real(tok, &p); pos.push_back(v); }
    else if (is_stmt(tok, "vn")) { p = 3; vec3 v; v.x = parse_real(tok, &p); v.y = parse_real(tok, &p); v.z = parse_real(tok, &p); nrm.push_back(v); }
    else if (is_stmt(tok, "vt")) { p = 3; vec2 v; v.x = parse_real(tok, &p); v.y = parse_real(tok, &p); tex.push_back(v); }
    else if (is_stmt(tok, "f")) {
      const size_t n = tok.size();
      const char* c = tok.c_str();
      bool bad = false;
      auto fix = [&](int idx, size_t count) { if (idx == 0) bad = true; return idx > 0 ? idx - 1 : (int)count + idx; };
      auto skip_index = [&](size_t q) { while (q < n && c[q] != '/' && !is_blank(c[q]) && c[q] != '\r') q++; return q; };
      std::vector<Corner> face;
      p = skip_blanks(tok, 2);
      while (p < n && !bad) {   // i, i/j, i//k, i/j/k
        Corner cr = {fix(std::atoi(c + p), pos.size()), -1, -1};
        p = skip_index(p);
        if (p < n && c[p] == '/') {
          p++;
          if (p < n && c[p] == '/') {
            p++;
            cr.n = fix(std::atoi(c + p), nrm.size());
            p = skip_index(p);
          } else {
            cr.t = fix(std::atoi(c + p), tex.size());
            p = skip_index(p);
            if (p < n && c[p] == '/') {
              p++;
              cr.n = fix(std::atoi(c + p), nrm.size());
              p = skip_index(p);
            }
          }
        }
        if (cr.t < -1) cr.t = -1;   // a relative normal / texcoord index in front of the array counts as "none" (load_model tests >= 0, renderer.cpp:489,499)
        if (cr.n < -1) cr.n = -1;
        face.push_back(cr);
        p = skip_blanks(tok, p);
      }
      if (bad) return false;   // "Failed parse `f' line(e.g. zero value for face index)": LoadObj fails
      // an element that does not exist is undefined behaviour in the reference (renderer.cpp:466-498 reads past the arrays): refused here
      for (const Corner& cr : face)
        if (cr.v < 0 || (size_t)cr.v >= pos.size() || (cr.t != -1 && (cr.t < 0 || (size_t)cr.t >= tex.size())) || (cr.n != -1 && (cr.n < 0 || (size_t)cr.n >= nrm.size())))
          return false;
      group.push_back(face);
    } else if (is_stmt(tok, "usemtl")) {
      const auto it = by_name.find(tok.substr(7));
      const int now = it == by_name.end() ? -1 : it->second;
      if (now != material) { export_group(); group.clear(); material = now; }
    } else if (is_stmt(tok, "mtllib")) {
      std::stringstream names(tok.substr(7));
      std::string name;
      bool found = false;
      while (!found && std::getline(names, name, ' ')) found = parse_mtl(base + name, &out->materials, &by_name);
      if (!found) std::fprintf(stderr, "load_obj: material file(s) of '%s' not found, default material used\n", tok.c_str());   // the loader warns and continues
    } else if (is_stmt(tok, "g")) {
      export_group();
      if (!shape.empty()) shapes.push_back(shape);
      shape.clear(); group.clear();
    } else if (is_stmt(tok, "o")) {
      if (export_group()) shapes.push_back(shape);
      shape.clear(); group.clear();
    }
  }
  if (export_group() || !shape.empty()) shapes.push_back(shape);

  // ---- Renderer::load_model: dedupe in first-use order, one range per shape and per run of one material
  std::unordered_map<VertexKey, uint32_t, VertexKeyHash> uniq;
  for (const std::vector<Tri>& sh : shapes) {
    size_t start = out->indices.size();
    int run_mat = 0;
    bool open = false;
    for (const Tri& t : sh) {
      if (open && t.material != run_mat) {
        out->ranges.push_back({(uint32_t)start, (uint32_t)(out->indices.size() - start), run_mat});
        start = out->indices.size();
      }
      run_mat = t.material; open = true;
      for (const Corner& c : t.c) {
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
    if (out->indices.size() > start) out->ranges.push_back({(uint32_t)start, (uint32_t)(out->indices.size() - start), run_mat});
  }
  return !out->indices.empty();
}

}  // namespace vct
