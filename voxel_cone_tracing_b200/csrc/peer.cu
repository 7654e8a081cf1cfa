// peer.cu -- multi-GPU exchange over NVLink peer memory (one process per GPU, one node).
//
// No reference counterpart (the reference is single-GPU).  Design (BASELINE.json north_star: z-slab
// voxelization + exchange of the base level, local mip build, screen-tile split of the trace), built on CUDA IPC
// instead of a collective library so that the exchange is FUSED into the producing kernels:
//   * every rank exports its two level-0 buffers (double buffered), its frame and a small flag block
//     (vct_peer_export), the handles are exchanged by the launcher (bench.py: torch.distributed all_gather_object)
//     and mapped with cudaIpcOpenMemHandle (vct_peer_connect) -> plain device pointers into the peers' HBM;
//   * vox_resolve_kernel stores every resolved voxel of its z-slab into all peers' grids (SPARSE: only occupied
//     voxels cross NVLink -- 0.3 % of the grid in the reference scene -- where an all-gather ships all of it);
//   * shade_kernel stores the finished pixels of its screen tiles into the frame of the root rank (or all ranks);
//   * ordering is by epoch flags: the last block of the pushing kernel publishes the frame number into the
//     destination's flag block (fence.sys + st.release.sys), the consumer spins on ld.acquire.sys in a one-warp
//     kernel in front of the mip build / at the end of the frame.  One flag wait per exchange, no host round trip.
// Buffer reuse is made safe by double buffering level 0 (see vct_render_frame in capi.cu for the proof sketch).
#include "vct_internal.cuh"

namespace vct {

struct PeerHandlePack {   // the payload of vct_peer_handle_t
  cudaIpcMemHandle_t base[2], frame, flags;
  uint32_t R, W, H, magic;
  uint32_t n_tiles;         // > 0: the flag block is followed by the mip tile flags of the two level-0 buffers (n_tiles bytes each)
};
static_assert(sizeof(PeerHandlePack) <= sizeof(vct_peer_handle_t), "vct_peer_handle_t too small");
constexpr uint32_t kPeerMagic = 0x56435450u;   // "VCTP"
constexpr size_t kFlagWords = PEER_FLAG_KINDS * VCT_MAX_RANKS + 8;   // flags + done counters [PEER_FLAG_KINDS] + error word

__global__ void peer_wait_kernel(const uint32_t* flags, int kind, int nranks, uint32_t epoch, uint32_t* error_word, uint32_t* status, long long timeout_cycles) {
  const int p = threadIdx.x;
  if (p >= nranks) return;
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(flags + kind * VCT_MAX_RANKS + p) - epoch) < 0) {
    if (clock64() - t0 > timeout_cycles) {   // a peer died or never connected: the next host call on this device returns an error
      atomicExch(error_word, 1u + (uint32_t)p);
      *reinterpret_cast<volatile uint32_t*>(status + STATUS_PEER) = 1u + (uint32_t)p;
      __threadfence_system();
      return;
    }
    __nanosleep(100);
  }
}

// Zero, on EVERY rank, the voxels this rank stored into this frame's level-0 buffer the last time it was in use (two frames ago), and
// un-mark their mip tiles: the sparse counterpart of every rank clearing 4 R^3 bytes.  A voxel belongs to the slab of exactly one rank,
// a tile (8 slices) to one slab, so no two ranks ever write the same word or flag.
// Counts live in a ring of four slots indexed by the frame number: frame e appends under slot e & 3, reads (here) the count of frame e - 2 and
// zeroes the slot of frame e + 1 -- no separate reset launch, and nobody reads a slot while it is being zeroed.
__global__ void __launch_bounds__(256)
peer_unpush_kernel(const PeerView pv, int logR, const uint32_t* old_count, uint32_t* next_count) {
  const uint32_t n = min(*old_count, pv.pushed_capacity);
  if (blockIdx.x == 0 && threadIdx.x == 0) *next_count = 0u;
  const uint32_t m = (1u << logR) - 1u;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t voxel = pv.pushed[i];
    const uint32_t x = voxel & m, y = (voxel >> logR) & m, z = voxel >> (2 * logR);
    const uint32_t tile = ((((z >> 3) << (logR - 3)) + (y >> 3)) << (logR - 5)) + (x >> 5);   // tile_of_voxel (voxelize.cu)
    for (int p = 0; p < pv.nranks; p++) {
      pv.base[p][voxel] = 0u;
      pv.touched[p][tile] = 0;
    }
  }
  __threadfence_system();   // performed on the peers before this rank's next flag goes out (the resolve kernel that follows publishes it)
}
int launch_peer_unpush(vct_device* dev, const PeerView& pv, int logR) {
  const uint32_t e = pv.epoch;
  peer_unpush_kernel<<<dev->prop.multiProcessorCount * 2, 256, 0, dev->stream>>>(pv, logR, dev->pushed_n + ((e + 2u) & 3u), dev->pushed_n + ((e + 1u) & 3u));
  VCT_CUDA(cudaGetLastError());
  return VCT_OK;
}

// the lists hold at most one entry per fragment slot; they follow the arena when it grows (content kept: an un-push may be pending)
int ensure_pushed_lists(vct_device* dev) {
  if (dev->pushed_capacity >= dev->frag_capacity && dev->pushed_n) return VCT_OK;
  const size_t cap = dev->frag_capacity;
  if (!dev->pushed_n) {
    VCT_CUDA(cudaMalloc(&dev->pushed_n, 4 * sizeof(uint32_t)));
    VCT_CUDA(cudaMemsetAsync(dev->pushed_n, 0, 4 * sizeof(uint32_t), dev->stream));
  }
  for (int b = 0; b < 2; b++) {
    uint32_t* nl = nullptr;
    VCT_CUDA(cudaMalloc(&nl, cap * sizeof(uint32_t)));
    if (dev->pushed_list[b]) {
      VCT_CUDA(cudaMemcpyAsync(nl, dev->pushed_list[b], dev->pushed_capacity * sizeof(uint32_t), cudaMemcpyDeviceToDevice, dev->stream));
      VCT_CUDA(cudaStreamSynchronize(dev->stream));
      cudaFree(dev->pushed_list[b]);
    }
    dev->pushed_list[b] = nl;
  }
  dev->pushed_capacity = cap;
  return VCT_OK;
}

// a flag of its own launch: "everything this device object had queued before has completed" (replicated voxelization has no
// pushing kernel to carry the PUSHED flag)
__global__ void peer_signal_kernel(const PeerView pv, int kind) { peer_signal_last_block(pv, kind, -1); }
int launch_peer_signal(vct_device* dev, const PeerView& pv, int kind) {
  peer_signal_kernel<<<1, 32, 0, dev->stream>>>(pv, kind);
  VCT_CUDA(cudaGetLastError());
  return VCT_OK;
}

int launch_peer_wait(vct_device* dev, int kind, uint32_t epoch) {
  const long long timeout = (long long)dev->prop.clockRate * 1000ll * 5ll;   // ~5 s of SM clock (clockRate is in kHz)
  peer_wait_kernel<<<1, 32, 0, dev->stream>>>(dev->peer_flags, kind, dev->peers.nranks, epoch, dev->peer_flags + kFlagWords - 1, dev->status_dev, timeout);
  VCT_CUDA(cudaGetLastError());
  return VCT_OK;
}

}  // namespace vct

using namespace vct;

extern "C" {

int vct_peer_export(vct_device_t* dev, vct_grid_t* g, vct_target_t* t, vct_peer_handle_t* out) {
  VCT_REQUIRE(dev && g && t && out, "null argument");
  VCT_REQUIRE(g->fmt == VCT_GRID_RGBA8, "the multi-GPU exchange handles RGBA8 grids only");
  VCT_CUDA(cudaSetDevice(dev->ordinal));
  const size_t n0 = (size_t)g->R * g->R * g->R * 4;
  if (!g->base_buf[1]) {
    VCT_CUDA(cudaMalloc(&g->base_buf[1], n0));
    g->bytes += n0;
  }
  g->external = true;   // peers store into level 0: the sparse clear / sparse mip bookkeeping does not apply
  g->untrack();
  VCT_CUDA(cudaMemsetAsync(g->base_buf[0], 0, n0, dev->stream));
  VCT_CUDA(cudaMemsetAsync(g->base_buf[1], 0, n0, dev->stream));
  // the flag block, followed (grids the streaming mip kernel handles) by the mip tile flags of the two level-0 buffers: the peers mark them
  const size_t n_tiles = g->tile_touched ? (size_t)g->R * g->R * g->R / 2048 : 0;
  const size_t flag_bytes = kFlagWords * 4 + 2 * n_tiles;
  if (dev->peer_flags && dev->peer_flag_bytes != flag_bytes) { cudaFree(dev->peer_flags); dev->peer_flags = nullptr; }
  if (!dev->peer_flags) VCT_CUDA(cudaMalloc(&dev->peer_flags, flag_bytes));
  dev->peer_flag_bytes = flag_bytes;
  VCT_CUDA(cudaMemsetAsync(dev->peer_flags, 0, flag_bytes, dev->stream));
  if (dev->pushed_n) VCT_CUDA(cudaMemsetAsync(dev->pushed_n, 0, 4 * sizeof(uint32_t), dev->stream));   // both buffers are zero again: nothing to un-push
  VCT_CUDA(cudaStreamSynchronize(dev->stream));
  PeerHandlePack h;
  memset(&h, 0, sizeof h);
  VCT_CUDA(cudaIpcGetMemHandle(&h.base[0], g->base_buf[0]));
  VCT_CUDA(cudaIpcGetMemHandle(&h.base[1], g->base_buf[1]));
  VCT_CUDA(cudaIpcGetMemHandle(&h.frame, t->frame));
  VCT_CUDA(cudaIpcGetMemHandle(&h.flags, dev->peer_flags));
  h.R = (uint32_t)g->R; h.W = (uint32_t)t->W; h.H = (uint32_t)t->H; h.magic = kPeerMagic;
  h.n_tiles = (uint32_t)n_tiles;
  dev->peer_export_fresh = true;   // the flag block is zero: epochs restart at 1 with the next connect
  memset(out, 0, sizeof *out);
  memcpy(out, &h, sizeof h);
  return VCT_OK;
}

int vct_peer_disconnect(vct_device_t* dev) {
  VCT_REQUIRE(dev, "device is null");
  cudaSetDevice(dev->ordinal);
  cudaStreamSynchronize(dev->stream);
  for (int i = 0; i < dev->n_peer_mapped; i++) cudaIpcCloseMemHandle(dev->peer_mapped[i]);
  dev->n_peer_mapped = 0;
  if (dev->peer_grid) { dev->peer_grid->base = dev->peer_grid->base_buf[0]; dev->peer_grid->peer_touched = nullptr; }
  memset(dev->peer_touched_all, 0, sizeof dev->peer_touched_all);
  memset(&dev->peers, 0, sizeof dev->peers);
  dev->peer_grid = nullptr; dev->peer_target = nullptr;
  dev->peer_epoch = 0;
  dev->peer_sparse_mode = -1;
  dev->peer_replicate = -1;
  return VCT_OK;
}

int vct_peer_connect(vct_device_t* dev, vct_grid_t* g, vct_target_t* t, int rank, int nranks, const vct_peer_handle_t* all, int frame_root) {
  VCT_REQUIRE(dev && g && t && all, "null argument");
  VCT_REQUIRE(nranks >= 1 && nranks <= VCT_MAX_RANKS && rank >= 0 && rank < nranks, "bad rank / nranks");
  VCT_REQUIRE(frame_root >= -1 && frame_root < nranks, "bad frame_root");
  VCT_REQUIRE(g->base_buf[1] && dev->peer_flags, "call vct_peer_export first");
  // a connect restarts the epochs at 1; flags left over from an earlier connection would let the first waits pass at once
  VCT_REQUIRE(dev->peer_export_fresh, "vct_peer_connect needs a fresh vct_peer_export (it zeroes the flag block) on every rank");
  dev->peer_export_fresh = false;
  VCT_REQUIRE(g->R >= nranks, "more ranks than z-slices");
  VCT_CUDA(cudaSetDevice(dev->ordinal));
  vct_peer_disconnect(dev);
  PeerView pv;
  memset(&pv, 0, sizeof pv);
  pv.rank = rank; pv.nranks = nranks; pv.frame_root = frame_root;
  pv.done_counter = dev->peer_flags + PEER_FLAG_KINDS * VCT_MAX_RANKS;
  for (int p = 0; p < nranks; p++) {
    PeerHandlePack h;
    memcpy(&h, &all[p], sizeof h);
    VCT_REQUIRE(h.magic == kPeerMagic, "peer handle is not a vct_peer_handle_t");
    VCT_REQUIRE(h.R == (uint32_t)g->R && h.W == (uint32_t)t->W && h.H == (uint32_t)t->H, "peer grid / frame size differs from the local one");
    if (p == rank) {
      dev->peer_base_all[0][p] = g->base_buf[0]; dev->peer_base_all[1][p] = g->base_buf[1];
      pv.frame[p] = t->frame; pv.flags[p] = dev->peer_flags;
      for (int b = 0; b < 2; b++) dev->peer_touched_all[b][p] = h.n_tiles ? (uint8_t*)(dev->peer_flags + kFlagWords) + (size_t)b * h.n_tiles : nullptr;
      continue;
    }
    void* ptr[4] = {nullptr, nullptr, nullptr, nullptr};
    const cudaIpcMemHandle_t hs[4] = {h.base[0], h.base[1], h.frame, h.flags};
    for (int k = 0; k < 4; k++) {
      cudaError_t e = cudaIpcOpenMemHandle(&ptr[k], hs[k], cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        set_error("cudaIpcOpenMemHandle(rank %d, buffer %d) failed: %s", p, k, cudaGetErrorString(e));
        vct_peer_disconnect(dev);
        return VCT_ERR_CUDA;
      }
      dev->peer_mapped[dev->n_peer_mapped++] = ptr[k];
    }
    dev->peer_base_all[0][p] = (uint32_t*)ptr[0]; dev->peer_base_all[1][p] = (uint32_t*)ptr[1];
    pv.frame[p] = (uint32_t*)ptr[2]; pv.flags[p] = (uint32_t*)ptr[3];
    for (int b = 0; b < 2; b++) dev->peer_touched_all[b][p] = h.n_tiles ? (uint8_t*)((uint32_t*)ptr[3] + kFlagWords) + (size_t)b * h.n_tiles : nullptr;
  }
  dev->peers = pv;
  dev->peer_grid = g; dev->peer_target = t;
  dev->peer_epoch = 0;
  return VCT_OK;
}

int vct_peer_error(vct_device_t* dev) {
  VCT_REQUIRE(dev, "device is null");
  if (!dev->peer_flags) return VCT_OK;
  uint32_t w = 0;
  VCT_CUDA(cudaMemcpyAsync(&w, dev->peer_flags + kFlagWords - 1, 4, cudaMemcpyDeviceToHost, dev->stream));
  VCT_CUDA(cudaStreamSynchronize(dev->stream));
  if (w) {
    set_error("peer wait timed out: rank %u never signalled (peer process dead, not connected, or frames out of step)", w - 1u);
    return VCT_ERR_CUDA;
  }
  return VCT_OK;
}

}  // extern "C"
