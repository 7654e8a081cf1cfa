// Host build (g++) of the mip arithmetic header the CUDA kernels use (csrc/mip_arith.cuh): lets the CPU test suite check the
// integer dot-product formulation + tie replay against the oracle's fp32 recipe without a GPU.  Test infrastructure only.
#include <stdint.h>
#include <stddef.h>

#include "mip_arith.cuh"

using namespace vct;

extern "C" {

// One mip step.  shared != 0: src is ONE volume (Ns^3 words) used by all six directions (level 0 -> 1);
// otherwise src holds six volumes (direction-major).  dst: six volumes of (Ns/2)^3.  Returns the number of replayed ties.
uint64_t vct_hosttest_mip_step(const uint32_t* src, int Ns, int shared, uint32_t* dst) {
  const int Nd = Ns / 2;
  const size_t ns = (size_t)Ns * Ns * Ns, nd = (size_t)Nd * Nd * Nd;
  uint64_t n_ties = 0;
  for (int z = 0; z < Nd; z++)
    for (int y = 0; y < Nd; y++)
      for (int x = 0; x < Nd; x++) {
        const size_t o = ((size_t)z * Nd + y) * Nd + x;
        if (shared) {
          uint32_t w[8];
          for (int dz = 0; dz < 2; dz++)
            for (int dy = 0; dy < 2; dy++)
              for (int dx = 0; dx < 2; dx++) w[child_id(dx, dy, dz)] = src[((size_t)(2 * z + dz) * Ns + (2 * y + dy)) * Ns + (2 * x + dx)];
          uint32_t out[6], ties;
          mip_filter6_shared(w, out, ties);
          for (; ties; ties &= ties - 1) {
            const int bit = __builtin_ctz(ties), d = bit >> 2, k = bit & 3;
            const uint32_t b = mip_replay_channel([&](int i) { return w[i]; }, d, k);
            out[d] = (out[d] & ~(0xFFu << (8 * k))) | (b << (8 * k));
            n_ties++;
          }
          for (int d = 0; d < 6; d++) dst[d * nd + o] = out[d];
        } else {
          for (int d = 0; d < 6; d++) {
            uint32_t w[8];
            for (int dz = 0; dz < 2; dz++)
              for (int dy = 0; dy < 2; dy++)
                for (int dx = 0; dx < 2; dx++) w[child_id(dx, dy, dz)] = src[d * ns + ((size_t)(2 * z + dz) * Ns + (2 * y + dy)) * Ns + (2 * x + dx)];
            uint32_t ties;
            uint32_t out = mip_filter1(w, d, ties);
            for (; ties; ties &= ties - 1) {
              const int k = __builtin_ctz(ties);
              const uint32_t b = mip_replay_channel([&](int i) { return w[i]; }, d, k);
              out = (out & ~(0xFFu << (8 * k))) | (b << (8 * k));
              n_ties++;
            }
            dst[d * nd + o] = out;
          }
        }
      }
  return n_ties;
}

}  // extern "C"

// tie statistics of a level-0 -> 1 step as the fused kernel sees them: per warp (16 x 2 texels of a 32x16x8 tile's level-1 plane) the
// largest number of tied channels of any lane.  hist[0..24] counts warps that have at least one non-zero texel.
extern "C" void vct_hosttest_tie_hist(const uint32_t* src, int Ns, uint64_t* hist) {
  using namespace vct;
  const int Nd = Ns / 2;
  for (int i = 0; i <= 24; i++) hist[i] = 0;
  for (int z = 0; z < Nd; z++)
    for (int y0 = 0; y0 < Nd; y0 += 2)
      for (int x0 = 0; x0 < Nd; x0 += 16) {
        int mx = 0; bool any_nz = false;
        for (int y = y0; y < y0 + 2; y++)
          for (int x = x0; x < x0 + 16; x++) {
            uint32_t w[8], any = 0;
            for (int dz = 0; dz < 2; dz++)
              for (int dy = 0; dy < 2; dy++)
                for (int dx = 0; dx < 2; dx++) { w[child_id(dx, dy, dz)] = src[((size_t)(2 * z + dz) * Ns + (2 * y + dy)) * Ns + (2 * x + dx)]; any |= w[child_id(dx, dy, dz)]; }
            if (!any) continue;
            any_nz = true;
            uint32_t out[6], ties;
            mip_filter6_shared(w, out, ties);
            int pc = __builtin_popcount(ties);
            if (pc > mx) mx = pc;
          }
        if (any_nz) hist[mx]++;
      }
}
