// texture_3d.h -- the five 3-D texture helpers of the reference (src/texture_3d.h:6-10) on CUDA memory.
//
// Reference semantics kept (src/texture_3d.cpp:3-51): RGBA8 unorm storage with `levels` mip levels
// (glTexStorage3D: levels must not exceed floor(log2(max dim)) + 1), clear = fill level 0 with one colour,
// mip_tex_3d = isotropic 2x2x2 box filter of every level (glGenerateMipmap; not used by Renderer, which
// builds the six directional chains with its own kernel).  Handles are opaque pointers instead of GLuint
// names; `activate_tex_3d` (bind to a texture unit of a GL program) has nothing left to bind and only
// validates the handle.  upload/download are additions (glTexSubImage3D / glGetTexImage equivalents).
#pragma once

#include <cstdint>

#include "vct/device.h"

namespace vct {

typedef vct_tex3d_t* tex3d_handle_t;   // replaces GLuint

tex3d_handle_t create_tex_3d(Device& device, int width, int height, int depth, int levels);  // nullptr on failure
void destroy_tex_3d(tex3d_handle_t tex);
void activate_tex_3d(unsigned program, tex3d_handle_t tex, unsigned unit);
void clear_tex_3d(tex3d_handle_t tex, float clear_color[4]);
void mip_tex_3d(tex3d_handle_t tex);
bool upload_tex_3d(tex3d_handle_t tex, int level, const uint32_t* rgba8);
bool download_tex_3d(tex3d_handle_t tex, int level, uint32_t* rgba8);

}  // namespace vct
