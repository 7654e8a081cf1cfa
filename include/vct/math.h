// math.h -- the few vector / matrix helpers the host layer needs (the reference uses vendored glm 0.9.9
// for these; the B200 build has no third-party dependency).  Layouts are glm's: vec3 = 3 floats, mat4 =
// 16 floats column-major, so a glm::mat4 can be passed wherever a vct::mat4 is expected (templates in
// renderer.h accept any 64-byte matrix type).  Semantics follow glm 0.9.9 as the reference calls it:
// perspective() takes fovy in RADIANS (the reference passes 45.0f, src/camera.h:23 -- kept), right-handed,
// depth -1..1; look_at = glm::lookAt RH.  Arithmetic is plain float in the same operation order as glm so
// that the matrices equal the oracle's (tests compare them bit for bit).
#pragma once

#include <cmath>
#include <cstring>

namespace vct {

struct vec2 { float x, y; };
struct vec3 { float x, y, z; };
struct vec4 { float x, y, z, w; };
struct mat4 {
  float m[16];  // column-major: m[4*c + r]
  static mat4 identity() { mat4 r; std::memset(r.m, 0, sizeof r.m); r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f; return r; }
};

inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(vec3 a, vec3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline vec3 normalize(vec3 a) { float s = 1.0f / std::sqrt(dot(a, a)); return {a.x * s, a.y * s, a.z * s}; }  // glm: v * inversesqrt(dot)

mat4 perspective(float fovy_radians, float aspect, float z_near, float z_far);  // glm::perspective (RH, NO)
mat4 look_at(vec3 eye, vec3 center, vec3 up);                                   // glm::lookAt (RH)
mat4 mul(const mat4& a, const mat4& b);                                         // a * b
mat4 translate(const mat4& m, vec3 v);                                          // glm::translate(m, v)
mat4 rotate(const mat4& m, float angle_radians, vec3 axis);                     // glm::rotate(m, angle, axis)
mat4 scale(const mat4& m, vec3 v);                                              // glm::scale(m, v)
vec3 camera_front(float pitch_degrees, float yaw_degrees);                      // Camera::calc_front, src/camera.h:25-37

}  // namespace vct
