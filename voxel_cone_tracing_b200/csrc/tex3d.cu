// tex3d.cu -- the generic single 3-D RGBA8 texture of src/texture_3d.h:6-10 (create / destroy /
// clear / isotropic mip).  The renderer's six directional voxel textures do NOT use this object
// (they are the deduplicated vct_grid); it exists so that the texture_3d.h surface is complete.
#include <new>

#include "vct_internal.cuh"

namespace vct {

// glGenerateMipmap equivalent for RGBA8: 2x2x2 box filter per channel, round half up.
__global__ void tex3d_box_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int sw, int sh, int sd, int dw, int dh, int dd) {
  const size_t n = (size_t)dw * dh * dd;
  for (size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x; u < n; u += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(u % dw), y = (int)((u / dw) % dh), z = (int)(u / ((size_t)dw * dh));
    uint32_t sum[4] = {0, 0, 0, 0}, cnt = 0;
    for (int dz = 0; dz < 2; dz++)
      for (int dy = 0; dy < 2; dy++)
        for (int dx = 0; dx < 2; dx++) {
          const int sx = min(2 * x + dx, sw - 1), sy = min(2 * y + dy, sh - 1), sz = min(2 * z + dz, sd - 1);
          const uint32_t w = src[((size_t)sz * sh + sy) * sw + sx];
          sum[0] += w & 0xFFu; sum[1] += (w >> 8) & 0xFFu; sum[2] += (w >> 16) & 0xFFu; sum[3] += w >> 24;
          cnt++;
        }
    uint32_t o = 0;
    for (int k = 0; k < 4; k++) o |= ((sum[k] + cnt / 2) / cnt) << (8 * k);
    dst[u] = o;
  }
}

int launch_tex3d_mip(vct_tex3d* t) {
  cudaStream_t s = t->dev->stream;
  for (int l = 0; l + 1 < t->levels; l++) {
    const int sw = max(t->w >> l, 1), sh = max(t->h >> l, 1), sd = max(t->d >> l, 1);
    const int dw = max(t->w >> (l + 1), 1), dh = max(t->h >> (l + 1), 1), dd = max(t->d >> (l + 1), 1);
    const size_t n = (size_t)dw * dh * dd;
    tex3d_box_kernel<<<grid_for(n), 256, 0, s>>>(t->lvl[l], t->lvl[l + 1], sw, sh, sd, dw, dh, dd);
  }
  VCT_CUDA(cudaGetLastError());
  return VCT_OK;
}

}  // namespace vct

using namespace vct;

static size_t level_texels(const vct_tex3d* t, int l) {
  return (size_t)max(t->w >> l, 1) * (size_t)max(t->h >> l, 1) * (size_t)max(t->d >> l, 1);
}

extern "C" {

int vct_tex3d_create(vct_device_t* dev, int w, int h, int d, int levels, vct_tex3d_t** out) {
  VCT_REQUIRE(dev && out, "null argument");
  VCT_REQUIRE(w > 0 && h > 0 && d > 0 && w <= 2048 && h <= 2048 && d <= 2048, "bad size");
  int max_levels = 1;
  for (int m = max(w, max(h, d)); m > 1; m >>= 1) max_levels++;
  VCT_REQUIRE(levels >= 1 && levels <= max_levels && levels <= VCT_MAX_LEVELS, "bad level count");  // glTexStorage3D rule
  vct_tex3d* t = new (std::nothrow) vct_tex3d();
  if (!t) { set_error("out of host memory"); return VCT_ERR_OOM; }
  t->dev = dev; t->w = w; t->h = h; t->d = d; t->levels = levels;
  for (int l = 0; l < levels; l++) {
    cudaError_t e = cudaMalloc(&t->lvl[l], level_texels(t, l) * 4);
    if (e != cudaSuccess) { set_error("tex3d allocation failed: %s", cudaGetErrorString(e)); vct_tex3d_destroy(t); return VCT_ERR_OOM; }
    cudaMemsetAsync(t->lvl[l], 0, level_texels(t, l) * 4, dev->stream);
  }
  *out = t;
  return VCT_OK;
}

int vct_tex3d_destroy(vct_tex3d_t* t) {
  if (!t) return VCT_OK;
  cudaStreamSynchronize(t->dev->stream);
  for (int l = 0; l < VCT_MAX_LEVELS; l++) cudaFree(t->lvl[l]);
  delete t;
  return VCT_OK;
}

int vct_tex3d_clear(vct_tex3d_t* t, const float c[4]) {
  VCT_REQUIRE(t && c, "null argument");
  uint32_t v = 0;
  for (int k = 0; k < 4; k++) v |= ((uint32_t)rintf(fminf(fmaxf(c[k], 0.f), 1.f) * 255.0f)) << (8 * k);
  return launch_fill_u32(t->dev->stream, t->lvl[0], level_texels(t, 0), v);
}

int vct_tex3d_mip(vct_tex3d_t* t) {
  VCT_REQUIRE(t, "texture is null");
  return launch_tex3d_mip(t);
}

int vct_tex3d_upload(vct_tex3d_t* t, int level, const uint32_t* host) {
  VCT_REQUIRE(t && host && level >= 0 && level < t->levels, "bad argument");
  VCT_CUDA(cudaMemcpyAsync(t->lvl[level], host, level_texels(t, level) * 4, cudaMemcpyHostToDevice, t->dev->stream));
  VCT_CUDA(cudaStreamSynchronize(t->dev->stream));
  return VCT_OK;
}

int vct_tex3d_download(vct_tex3d_t* t, int level, uint32_t* host) {
  VCT_REQUIRE(t && host && level >= 0 && level < t->levels, "bad argument");
  VCT_CUDA(cudaMemcpyAsync(host, t->lvl[level], level_texels(t, level) * 4, cudaMemcpyDeviceToHost, t->dev->stream));
  VCT_CUDA(cudaStreamSynchronize(t->dev->stream));
  return VCT_OK;
}

}  // extern "C"
