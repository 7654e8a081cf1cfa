// texture_3d.cpp -- the reference's 3-D texture helpers (src/texture_3d.cpp:3-51) over vct_tex3d_*.
#include "vct/texture_3d.h"

#include <cstdio>

namespace vct {

tex3d_handle_t create_tex_3d(Device& device, int width, int height, int depth, int levels) {
  if (!device.ok()) { std::fprintf(stderr, "create_tex_3d: no device\n"); return nullptr; }
  vct_tex3d_t* t = nullptr;
  if (vct_tex3d_create(device.handle(), width, height, depth, levels, &t) != VCT_OK) {
    std::fprintf(stderr, "create_tex_3d: %s\n", vct_last_error());   // GL would raise GL_INVALID_OPERATION and return a dead name
    return nullptr;
  }
  return t;
}

void destroy_tex_3d(tex3d_handle_t tex) { if (tex) vct_tex3d_destroy(tex); }

void activate_tex_3d(unsigned /*program*/, tex3d_handle_t tex, unsigned unit) {
  // glActiveTexture + glBindTexture (texture_3d.cpp:32-37): nothing to bind on CUDA; keep the unit range check
  if (!tex || unit >= MAX_BOUND_TEXTURES) std::fprintf(stderr, "activate_tex_3d: bad texture or unit %u\n", unit);
}

void clear_tex_3d(tex3d_handle_t tex, float clear_color[4]) {
  if (tex && vct_tex3d_clear(tex, clear_color) != VCT_OK) std::fprintf(stderr, "clear_tex_3d: %s\n", vct_last_error());
}

void mip_tex_3d(tex3d_handle_t tex) {
  if (tex && vct_tex3d_mip(tex) != VCT_OK) std::fprintf(stderr, "mip_tex_3d: %s\n", vct_last_error());
}

bool upload_tex_3d(tex3d_handle_t tex, int level, const uint32_t* rgba8) { return tex && vct_tex3d_upload(tex, level, rgba8) == VCT_OK; }
bool download_tex_3d(tex3d_handle_t tex, int level, uint32_t* rgba8) { return tex && vct_tex3d_download(tex, level, rgba8) == VCT_OK; }

}  // namespace vct
