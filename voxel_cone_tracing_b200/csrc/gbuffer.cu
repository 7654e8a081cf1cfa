// gbuffer.cu -- camera visibility pass producing the G-buffer the cone tracer shades.
//
// Replaces voxel_cone_tracing.vert + the fixed-function raster / GL_LESS depth test of
// Renderer::visualize() (src/renderer.cpp:355-390).  The reference is a forward renderer that
// shades every overdrawn fragment; the last writer of a pixel is the GL_LESS winner, so shading
// only that fragment gives the same image.  Three launches:
//   cam_setup_kernel    triangle-parallel vertex shader + projection + snapping + item count
//   cam_raster_kernel   one warp per 8x8 item; 64-bit atomicMin of (depth bits << 32 | triangle
//                       sequence) = GL_LESS with "first drawn wins ties"
//   cam_resolve_kernel  per pixel: perspective-correct world position and (un-renormalised)
//                       normal of the winning triangle (voxel_cone_tracing.vert:22-28)
// Built with -fmad=false; same evaluation order as the oracle (rules R1-R3, R8).
#include "raster.cuh"

namespace vct {

struct Mat4 { float m[16]; };
constexpr int kSmallCamPixels = 36;

__global__ void __launch_bounds__(kSetupThreads)
cam_setup_kernel(const vct_vertex_t* __restrict__ verts, const uint32_t* __restrict__ indices, const DrawRec* __restrict__ draws,
                 uint32_t n_draws, uint32_t n_tris, Mat4 pv, int W, int H, CamTri* __restrict__ out, uint32_t* __restrict__ item_local,
                 uint32_t* __restrict__ item_block, unsigned long long* __restrict__ vis, int tile_rank, int tile_nranks, int small_limit,
                 uint32_t* scan_ticket, uint32_t* scan_total) {
  uint32_t t = blockIdx.x * kSetupThreads + threadIdx.x;
  uint32_t count = 0;
  RasterTri srt;   // copy for the small-triangle path below
  srt.sign = 0; srt.imin = 0; srt.imax = -1; srt.jmin = 0; srt.jmax = -1;
  float sz0 = 0.f, sz1 = 0.f, sz2 = 0.f;
  if (t < n_tris) {
    const DrawRec& d = draws[find_draw(t, draws, n_draws)];
    uint32_t first = d.first_index + 3u * (t - d.tri_base);
    CamTri v;
    const float* m = d.model;
    const float* p = pv.m;
    float xw[3], yw[3];
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const vct_vertex_t vx = verts[d.vertex_base + indices[first + k]];
      float px = vx.pos[0], py = vx.pos[1], pz = vx.pos[2];
      float wx = ((m[0] * px + m[4] * py) + m[8] * pz) + m[12];
      float wy = ((m[1] * px + m[5] * py) + m[9] * pz) + m[13];
      float wz = ((m[2] * px + m[6] * py) + m[10] * pz) + m[14];
      float ww = ((m[3] * px + m[7] * py) + m[11] * pz) + m[15];
      v.world[k][0] = wx; v.world[k][1] = wy; v.world[k][2] = wz;
      const float* nm = d.nmat;
      float nx = (nm[0] * vx.norm[0] + nm[3] * vx.norm[1]) + nm[6] * vx.norm[2];
      float ny = (nm[1] * vx.norm[0] + nm[4] * vx.norm[1]) + nm[7] * vx.norm[2];
      float nz = (nm[2] * vx.norm[0] + nm[5] * vx.norm[1]) + nm[8] * vx.norm[2];
      float nl = sqrtf((nx * nx + ny * ny) + nz * nz);
      v.nn[k][0] = nx / nl; v.nn[k][1] = ny / nl; v.nn[k][2] = nz / nl;
      float cx = ((p[0] * wx + p[4] * wy) + p[8] * wz) + p[12] * ww;
      float cy = ((p[1] * wx + p[5] * wy) + p[9] * wz) + p[13] * ww;
      float cz = ((p[2] * wx + p[6] * wy) + p[10] * wz) + p[14] * ww;
      float cw = ((p[3] * wx + p[7] * wy) + p[11] * wz) + p[15] * ww;
      if (!(cw > 0.0f)) ok = false;
      float iw = 1.0f / cw;
      v.iw[k] = iw;
      xw[k] = (cx * iw + 1.0f) * ((float)W * 0.5f);
      yw[k] = (cy * iw + 1.0f) * ((float)H * 0.5f);
      v.zw[k] = (cz * iw + 1.0f) * 0.5f;
    }
    if (ok) raster_setup(xw, yw, W, H, v.rt);
    else { v.rt.sign = 0; v.rt.imin = 0; v.rt.imax = -1; v.rt.jmin = 0; v.rt.jmax = -1; v.rt.area = 0; }
    v.material = d.material;
    v.pad = 0;
    out[t] = v;   // always: the resolve kernel reads the record of whichever triangle wins a pixel
    count = raster_item_count(v.rt);
    srt = v.rt; sz0 = v.zw[0]; sz1 = v.zw[1]; sz2 = v.zw[2];
  }
  // ---- small triangles (bounding box of at most kSmallCamPixels pixel centres): depth-tested right here, one lane per
  //      triangle, instead of one warp per 8x8 item ----
  const int bw = srt.imax - srt.imin + 1, bh = srt.jmax - srt.jmin + 1;
  const bool small = count > 0 && bw * bh <= small_limit;
  const int npx = small ? bw * bh : 0;
  for (int p = 0; p < npx; p++) {
    const int i = srt.imin + p % bw, j = srt.jmin + p / bw;
    if (tile_nranks > 1 && ((j >> 5) * ((W + 31) >> 5) + (i >> 5)) % tile_nranks != tile_rank) continue;   // multi-GPU: not this rank's screen tile
    float b[3];
    if (raster_sample(srt, i, j, b)) {
      const float zw = interp3(b, sz0, sz1, sz2);
      if (zw >= 0.0f && zw <= 1.0f) atomicMin(&vis[(size_t)j * W + i], ((unsigned long long)__float_as_uint(zw) << 32) | (unsigned long long)t);
    }
  }
  if (small) count = 0;
  block_scan_items(count, t, n_tris, item_local, item_block, scan_ticket, scan_total);
}

__global__ void __launch_bounds__(256)
cam_raster_kernel(const CamTri* __restrict__ tris, uint32_t n_tris, const uint32_t* __restrict__ item_local,
                  const uint32_t* __restrict__ item_block, uint32_t n_blocks, int W, unsigned long long* __restrict__ vis,
                  const uint32_t* __restrict__ counters, int tile_rank, int tile_nranks) {
  const uint32_t total = counters[CNT_CAM_ITEMS];
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t g = warp; g < total; g += n_warps) {  // one warp per 8x8 item
    uint32_t rank;
    const uint32_t ti = find_item_triangle(g, item_block, n_blocks, item_local, n_tris, rank);
    const CamTri& v = tris[ti];
    const RasterTri rt = v.rt;
    const float z0 = v.zw[0], z1 = v.zw[1], z2 = v.zw[2];
    const int tiles_x = (rt.imax >> 3) - (rt.imin >> 3) + 1;
    const int tx = (rt.imin >> 3) + (int)(rank % (uint32_t)tiles_x), ty = (rt.jmin >> 3) + (int)(rank / (uint32_t)tiles_x);
    // multi-GPU: only the 32x32 screen tiles this rank shades need visibility
    if (tile_nranks > 1 && ((ty >> 2) * ((W + 31) >> 5) + (tx >> 2)) % tile_nranks != tile_rank) continue;
    EdgeBlock eb;
    edge_block_setup(rt, tx * kTile, ty * kTile, eb);
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int p = lane + 32 * h;
      const int i = tx * kTile + (p & 7), j = ty * kTile + (p >> 3);
      float b[3];
      if (i >= rt.imin && i <= rt.imax && j >= rt.jmin && j <= rt.jmax && edge_block_sample(eb, p & 7, p >> 3, b)) {
        const float zw = interp3(b, z0, z1, z2);
        if (zw >= 0.0f && zw <= 1.0f) {  // near / far (R3); also rejects NaN
          const unsigned long long key = ((unsigned long long)__float_as_uint(zw) << 32) | (unsigned long long)ti;
          atomicMin(&vis[(size_t)j * W + i], key);
        }
      }
    }
  }
}

// key of the cleared depth buffer: depth 1.0 (glClear), no triangle.  A fragment at zw == 1.0 has a
// smaller key only if its triangle id is smaller than 0xFFFFFFFF, so mask it explicitly below.
constexpr unsigned long long kVisClear = ((unsigned long long)0x3F800000u << 32) | 0xFFFFFFFFull;

__global__ void __launch_bounds__(256)
cam_resolve_kernel(const CamTri* __restrict__ tris, const unsigned long long* vis, int W, int H, float* __restrict__ world_pos,
                   float* __restrict__ normal, uint32_t* __restrict__ material, unsigned long long* vis_out, int tile_rank, int tile_nranks) {
  const size_t n = (size_t)W * H;
  for (size_t px = (size_t)blockIdx.x * blockDim.x + threadIdx.x; px < n; px += (size_t)gridDim.x * blockDim.x) {
    if (tile_nranks > 1) {  // pixels of other ranks' tiles are never read by this rank's tracer
      const int i = (int)(px % W), j = (int)(px / W);
      if (((j >> 5) * ((W + 31) >> 5) + (i >> 5)) % tile_nranks != tile_rank) continue;
    }
    unsigned long long key = vis[px];
    uint32_t ti = (uint32_t)(key & 0xFFFFFFFFull);
    // GL_LESS against the cleared depth 1.0: a fragment exactly at zw == 1.0 fails
    if ((uint32_t)(key >> 32) >= 0x3F800000u) ti = VCT_NO_TRIANGLE;
    if (ti == VCT_NO_TRIANGLE) {
      if (vis_out) vis_out[px] = kVisClear;
      material[px] = VCT_NO_TRIANGLE;
      continue;
    }
    const CamTri& v = tris[ti];
    const int i = (int)(px % W), j = (int)(px / W);
    float b[3];
    raster_sample(v.rt, i, j, b);
    float q[3] = {b[0] * v.iw[0], b[1] * v.iw[1], b[2] * v.iw[2]};
    float qs = (q[0] + q[1]) + q[2];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      world_pos[px * 3 + c] = interp3(q, v.world[0][c], v.world[1][c], v.world[2][c]) / qs;
      normal[px * 3 + c] = interp3(q, v.nn[0][c], v.nn[1][c], v.nn[2][c]) / qs;
    }
    material[px] = v.material;
  }
}

__global__ void fill_u64_kernel(unsigned long long* p, size_t n, unsigned long long v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void fill_u32_kernel(uint32_t* p, size_t n, uint32_t v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void copy_u32x4_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n4, uint32_t* dst_tail, const uint32_t* src_tail, int n_tail) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
  if (blockIdx.x == 0 && (int)threadIdx.x < n_tail) dst_tail[threadIdx.x] = src_tail[threadIdx.x];
}
// device-to-device copy of n words by a kernel (16-byte accesses): a cudaMemcpyAsync would go through a copy engine
int launch_copy_u32(cudaStream_t s, uint32_t* dst, const uint32_t* src, size_t n) {
  if (n == 0) return VCT_OK;
  const size_t n4 = n / 4;
  copy_u32x4_kernel<<<grid_for(n4 ? n4 : 1, 256, 148 * 8), 256, 0, s>>>((uint4*)dst, (const uint4*)src, n4, dst + 4 * n4, src + 4 * n4, (int)(n - 4 * n4));
  VCT_CUDA(cudaGetLastError());
  return VCT_OK;
}
int launch_fill_u32(cudaStream_t s, uint32_t* p, size_t n, uint32_t v) {
  if (n == 0) return VCT_OK;
  fill_u32_kernel<<<grid_for(n), 256, 0, s>>>(p, n, v);
  VCT_CUDA(cudaGetLastError());
  return VCT_OK;
}

static void mat4_mul_host(const float* a, const float* b, float* out) {  // column-major a*b, no contraction on the host
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 4; r++) {
      volatile float s0 = a[0 * 4 + r] * b[c * 4 + 0];
      volatile float s1 = a[1 * 4 + r] * b[c * 4 + 1];
      volatile float s2 = a[2 * 4 + r] * b[c * 4 + 2];
      volatile float s3 = a[3 * 4 + r] * b[c * 4 + 3];
      volatile float t = s0 + s1;
      t = t + s2;
      out[c * 4 + r] = t + s3;
    }
}

int launch_gbuffer(vct_device* dev, vct_scene* sc, const float* view, const float* proj, vct_target_t_* t, int tile_rank, int tile_nranks) {
  cudaStream_t s = dev->stream;
  const size_t npx = (size_t)t->W * t->H;
  fill_u64_kernel<<<dev->prop.multiProcessorCount * 8, 256, 0, s>>>(t->vis, npx, kVisClear);
  const int sms = dev->prop.multiProcessorCount;
  if (sc->n_tris) {
    int rc = ensure_tri_scratch(dev, 1, sc->n_tris, sizeof(CamTri));
    if (rc) return rc;
    Mat4 pv;
    mat4_mul_host(proj, view, pv.m);  // projection * view (voxel_cone_tracing.vert:25)
    const uint32_t n_blocks = (sc->n_tris + kSetupThreads - 1) / kSetupThreads;
    CamTri* tris = (CamTri*)dev->rs[1].tri_recs;
    cam_setup_kernel<<<n_blocks, kSetupThreads, 0, s>>>(sc->verts, sc->indices, sc->draws, sc->n_draws, sc->n_tris, pv, t->W, t->H, tris,
                                                        dev->rs[1].item_local, dev->rs[1].item_block, t->vis, tile_rank, tile_nranks,
                                                        sc->n_tris >= kSmallPathMinTris ? kSmallCamPixels : 0, dev->counters + CNT_TICKET_CAM,
                                                        dev->counters + CNT_CAM_ITEMS);
    cam_raster_kernel<<<sms * 8, 256, 0, s>>>(tris, sc->n_tris, dev->rs[1].item_local, dev->rs[1].item_block, n_blocks, t->W, t->vis, dev->counters, tile_rank, tile_nranks);
    cam_resolve_kernel<<<sms * 8, 256, 0, s>>>(tris, t->vis, t->W, t->H, t->world_pos, t->normal, t->material, t->vis, tile_rank, tile_nranks);
  } else {
    launch_fill_u32(s, t->material, npx, VCT_NO_TRIANGLE);
  }
  VCT_CUDA(cudaGetLastError());
  return VCT_OK;
}

}  // namespace vct
