// Micro-benchmark behind the design of mip_stream_kernel: what do the level-0 read pattern and the surface-store
// pattern of the mip build achieve on their own?   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o stream_patterns stream_patterns.cu
//   read A: warp-tiles of 32x8x8 voxels (64 rows of 128 B), cp.async ring two slabs ahead, 12 warps / SM   (what the kernel does)
//   read B: the same tiles, plain 16-byte loads, all 16 of a tile issued at once
//   read C: contiguous grid-stride 16-byte loads (the copy-kernel pattern)
//   read D: warp-tiles of 128x2x... : one slab = 4 rows of 512 B
//   store S: 16-byte surface stores of zeros to a 3-D array, tile pattern of level 1 (16x4x4 texels per warp-tile x 6 directions)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ void cp16(void* s, const void* g) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(s)), "l"(g) : "memory"); }
__device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void wait2() { asm volatile("cp.async.wait_group 2;" ::: "memory"); }

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) read_a(const uint32_t* base, int R, uint32_t n_tiles, uint32_t* out) {
  __shared__ uint4 ring[WARPS][3][4][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, q = lane & 7, r = lane >> 3;
  const uint32_t gw = blockIdx.x * WARPS + wid, nw = gridDim.x * WARPS;
  const int ltx = __ffs(R / 32) - 1, lty = __ffs(R / 8) - 1;
  uint32_t acc = 0, sc = 0;
  auto issue = [&](uint32_t tile, int z1, uint32_t slot) {
    const int x0 = (tile & ((1u << ltx) - 1)) * 32, y0 = ((tile >> ltx) & ((1u << lty) - 1)) * 8, z0 = (tile >> (ltx + lty)) * 8;
    const uint32_t* p = base + ((size_t)(z0 + 2 * z1) * R + (y0 + 2 * r)) * R + x0 + 4 * q;
    cp16(&ring[wid][slot][0][lane], p); cp16(&ring[wid][slot][1][lane], p + R); cp16(&ring[wid][slot][2][lane], p + (size_t)R * R); cp16(&ring[wid][slot][3][lane], p + (size_t)R * R + R);
  };
  int ahead = 0;
  for (uint32_t t = gw; t < n_tiles; t += nw) {
    if (!ahead) { issue(t, 0, sc % 3); commit(); issue(t, 1, (sc + 1) % 3); commit(); }
    ahead = 0;
    for (int z1 = 0; z1 < 4; z1++, sc++) {
      if (z1 < 2) issue(t, z1 + 2, (sc + 2) % 3);
      else if (t + nw < n_tiles) { issue(t + nw, z1 - 2, (sc + 2) % 3); ahead = 2; }
      commit(); wait2();
      for (int i = 0; i < 4; i++) { uint4 v = ring[wid][sc % 3][i][lane]; acc |= v.x | v.y | v.z | v.w; }
    }
  }
  if (acc == 0x12345678u) out[0] = acc;
}
__global__ void __launch_bounds__(256) read_b(const uint32_t* base, int R, uint32_t n_tiles, uint32_t* out) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, q = lane & 7, r = lane >> 3;
  const uint32_t gw = blockIdx.x * 8 + wid, nw = gridDim.x * 8;
  const int ltx = __ffs(R / 32) - 1, lty = __ffs(R / 8) - 1;
  uint32_t acc = 0;
  for (uint32_t tile = gw; tile < n_tiles; tile += nw) {
    const int x0 = (tile & ((1u << ltx) - 1)) * 32, y0 = ((tile >> ltx) & ((1u << lty) - 1)) * 8, z0 = (tile >> (ltx + lty)) * 8;
    uint4 v[16];
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __ldcs((const uint4*)(base + ((size_t)(z0 + (i >> 1)) * R + (y0 + 2 * r + (i & 1))) * R + x0 + 4 * q));
#pragma unroll
    for (int i = 0; i < 16; i++) acc |= v[i].x | v[i].y | v[i].z | v[i].w;
  }
  if (acc == 0x12345678u) out[0] = acc;
}
__global__ void __launch_bounds__(256) read_c(const uint4* base, size_t n4, uint32_t* out) {
  uint32_t acc = 0;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * st < n4; i += 4 * st) {
    uint4 a = __ldcs(base + i), b = __ldcs(base + i + st), c = __ldcs(base + i + 2 * st), d = __ldcs(base + i + 3 * st);
    acc |= a.x | b.y | c.z | d.w;
  }
  for (; i < n4; i += st) acc |= __ldcs(base + i).x;
  if (acc == 0x12345678u) out[0] = acc;
}
// wide slabs: a warp-tile is 128 x 8 x 8 voxels; one slab = rows (2 y) x (2 z) of 512 B: lane = 16-byte piece of the row
__global__ void __launch_bounds__(256) read_d(const uint32_t* base, int R, uint32_t n_tiles, uint32_t* out) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t gw = blockIdx.x * 8 + wid, nw = gridDim.x * 8;
  const int ltx = __ffs(R / 128) - 1, lty = __ffs(R / 8) - 1;
  uint32_t acc = 0;
  for (uint32_t tile = gw; tile < n_tiles; tile += nw) {
    const int x0 = (tile & ((1u << ltx) - 1)) * 128, y0 = ((tile >> ltx) & ((1u << lty) - 1)) * 8, z0 = (tile >> (ltx + lty)) * 8;
    for (int s = 0; s < 16; s += 4) {   // 4 slabs at a time = 16 rows of 512 B
      uint4 v[16];
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const int row = s * 4 + i, y = row & 7, z = row >> 3;
        v[i] = __ldcs((const uint4*)(base + ((size_t)(z0 + z) * R + (y0 + y)) * R + x0 + 4 * lane));
      }
#pragma unroll
      for (int i = 0; i < 16; i++) acc |= v[i].x | v[i].y | v[i].z | v[i].w;
    }
  }
  if (acc == 0x12345678u) out[0] = acc;
}
// level-1 store pattern: warp-tile = 16 x 4 x 4 texels x 6 directions stacked in z
__global__ void __launch_bounds__(256) store_s(cudaSurfaceObject_t surf, int N1, int pitch, uint32_t n_tiles) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t gw = blockIdx.x * 8 + wid, nw = gridDim.x * 8;
  const int ltx = __ffs(N1 / 16) - 1, lty = __ffs(N1 / 4) - 1;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  for (uint32_t tile = gw; tile < n_tiles; tile += nw) {
    const int x0 = (tile & ((1u << ltx) - 1)) * 16, y0 = ((tile >> ltx) & ((1u << lty) - 1)) * 4, z0 = (tile >> (ltx + lty)) * 4;
#pragma unroll
    for (int i = 0; i < 12; i++) {
      const int u = lane + 32 * i, xq = u & 3, y = (u >> 2) & 3, z = (u >> 4) & 3, d = u >> 6;
      surf3Dwrite(zero, surf, (x0 + 4 * xq) * 4, y0 + y, z0 + z + d * pitch);
    }
  }
}
int main(int argc, char** argv) {
  const int R = argc > 1 ? atoi(argv[1]) : 256;
  const size_t n = (size_t)R * R * R;
  uint32_t *base, *out;
  CK(cudaMalloc(&base, n * 4)); CK(cudaMalloc(&out, 4)); CK(cudaMemset(base, 0, n * 4));
  uint32_t* flush; const size_t fl = 256u << 20; CK(cudaMalloc(&flush, fl));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const uint32_t n_tiles = (uint32_t)(n / 2048);
  auto run = [&](const char* name, auto launch, double bytes) {
    float best = 1e9f;
    for (int it = 0; it < 6; it++) {
      cudaMemsetAsync(flush, it, fl);   // push the array out of L2
      cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (it && ms < best) best = ms;
    }
    printf("R=%d %-44s %8.1f us  %7.1f GB/s  (%s)\n", R, name, best * 1e3, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
  };
  run("A cp.async ring, 32x8x8 tiles, 12 warps/SM", [&] { read_a<6><<<148 * 2, 192>>>(base, R, n_tiles, out); }, n * 4.0);
  run("A cp.async ring, 32x8x8 tiles, 24 warps/SM", [&] { read_a<8><<<148 * 3, 256>>>(base, R, n_tiles, out); }, n * 4.0);
  run("A cp.async ring, 32x8x8 tiles, 32 warps/SM", [&] { read_a<8><<<148 * 4, 256>>>(base, R, n_tiles, out); }, n * 4.0);
  run("B ldg x16 per tile, 32x8x8 tiles, 16 warps/SM", [&] { read_b<<<148 * 2, 256>>>(base, R, n_tiles, out); }, n * 4.0);
  run("B ldg x16 per tile, 32x8x8 tiles, 32 warps/SM", [&] { read_b<<<148 * 4, 256>>>(base, R, n_tiles, out); }, n * 4.0);
  run("C contiguous ldg x4, 32 warps/SM", [&] { read_c<<<148 * 4, 256>>>((const uint4*)base, n / 4, out); }, n * 4.0);
  run("D ldg x16, 128x8x8 tiles (512 B rows), 16 w/SM", [&] { read_d<<<148 * 2, 256>>>(base, R, n_tiles / 4, out); }, n * 4.0);
  run("D ldg x16, 128x8x8 tiles (512 B rows), 32 w/SM", [&] { read_d<<<148 * 4, 256>>>(base, R, n_tiles / 4, out); }, n * 4.0);
  // level-1 array: N1^3 x 6 (+ pads)
  const int N1 = R / 2, pitch = N1 + 64;
  cudaArray_t arr; cudaChannelFormatDesc fmt = cudaCreateChannelDesc<uchar4>();
  CK(cudaMalloc3DArray(&arr, &fmt, make_cudaExtent(N1, N1, 6 * (size_t)pitch), cudaArraySurfaceLoadStore));
  cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
  cudaSurfaceObject_t surf; CK(cudaCreateSurfaceObject(&surf, &rd));
  const uint32_t n1_tiles = (uint32_t)((size_t)N1 * N1 * N1 / 256);
  run("S surface stores 16 B, level-1 tile pattern, 16 w/SM", [&] { store_s<<<148 * 2, 256>>>(surf, N1, pitch, n1_tiles); }, (double)N1 * N1 * N1 * 24.0);
  run("S surface stores 16 B, level-1 tile pattern, 32 w/SM", [&] { store_s<<<148 * 4, 256>>>(surf, N1, pitch, n1_tiles); }, (double)N1 * N1 * N1 * 24.0);
  return 0;
}
