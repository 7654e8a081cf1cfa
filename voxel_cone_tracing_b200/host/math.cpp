// math.cpp -- glm-0.9.9-compatible camera / transform helpers (see include/vct/math.h).
#include "vct/math.h"

namespace vct {

// v / sqrt(dot(v,v)): the operation order the oracle and the Python harness use
static vec3 norm_div(vec3 a) { float l = std::sqrt(dot(a, a)); return {a.x / l, a.y / l, a.z / l}; }

mat4 perspective(float fovy, float aspect, float zn, float zf) {
  // glm perspectiveRH_NO (thirdparty/glm/glm/gtc/matrix_transform.inl:343-356); fovy in radians
  const float t = std::tan(fovy / 2.0f);
  mat4 r;
  std::memset(r.m, 0, sizeof r.m);
  r.m[0] = 1.0f / (aspect * t);
  r.m[5] = 1.0f / t;
  r.m[10] = -(zf + zn) / (zf - zn);
  r.m[11] = -1.0f;
  r.m[14] = -(2.0f * zf * zn) / (zf - zn);
  return r;
}

mat4 look_at(vec3 eye, vec3 center, vec3 up) {
  const vec3 f = norm_div(center - eye);
  const vec3 s = norm_div(cross(f, up));
  const vec3 u = cross(s, f);
  mat4 r;
  std::memset(r.m, 0, sizeof r.m);
  r.m[0] = s.x; r.m[4] = s.y; r.m[8] = s.z;
  r.m[1] = u.x; r.m[5] = u.y; r.m[9] = u.z;
  r.m[2] = -f.x; r.m[6] = -f.y; r.m[10] = -f.z;
  r.m[12] = -dot(s, eye); r.m[13] = -dot(u, eye); r.m[14] = dot(f, eye); r.m[15] = 1.0f;
  return r;
}

mat4 mul(const mat4& a, const mat4& b) {
  mat4 r;
  for (int c = 0; c < 4; c++)
    for (int row = 0; row < 4; row++)
      r.m[4 * c + row] = a.m[row] * b.m[4 * c] + a.m[4 + row] * b.m[4 * c + 1] + a.m[8 + row] * b.m[4 * c + 2] + a.m[12 + row] * b.m[4 * c + 3];
  return r;
}

mat4 translate(const mat4& m, vec3 v) {
  mat4 r = m;
  for (int row = 0; row < 4; row++) r.m[12 + row] = m.m[row] * v.x + m.m[4 + row] * v.y + m.m[8 + row] * v.z + m.m[12 + row];
  return r;
}

mat4 rotate(const mat4& m, float angle, vec3 axis_in) {
  const float c = std::cos(angle), s = std::sin(angle);
  const vec3 a = norm_div(axis_in);
  const vec3 t = {(1.0f - c) * a.x, (1.0f - c) * a.y, (1.0f - c) * a.z};
  float R[3][3];  // R[col][row]
  R[0][0] = c + t.x * a.x;       R[0][1] = t.x * a.y + s * a.z; R[0][2] = t.x * a.z - s * a.y;
  R[1][0] = t.y * a.x - s * a.z; R[1][1] = c + t.y * a.y;       R[1][2] = t.y * a.z + s * a.x;
  R[2][0] = t.z * a.x + s * a.y; R[2][1] = t.z * a.y - s * a.x; R[2][2] = c + t.z * a.z;
  mat4 r;
  for (int col = 0; col < 3; col++)
    for (int row = 0; row < 4; row++) r.m[4 * col + row] = m.m[row] * R[col][0] + m.m[4 + row] * R[col][1] + m.m[8 + row] * R[col][2];
  for (int row = 0; row < 4; row++) r.m[12 + row] = m.m[12 + row];
  return r;
}

mat4 scale(const mat4& m, vec3 v) {
  mat4 r = m;
  for (int row = 0; row < 4; row++) { r.m[row] = m.m[row] * v.x; r.m[4 + row] = m.m[4 + row] * v.y; r.m[8 + row] = m.m[8 + row] * v.z; }
  return r;
}

vec3 camera_front(float pitch_deg, float yaw_deg) {
  const float k = 0.01745329251994329576923690768489f;  // glm::radians
  const float p = pitch_deg * k, y = yaw_deg * k;
  return norm_div({std::cos(p) * std::cos(y), std::sin(p), std::cos(p) * std::sin(y)});
}

}  // namespace vct
