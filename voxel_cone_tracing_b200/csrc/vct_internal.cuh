// vct_internal.cuh -- objects behind the opaque handles of include/vct/vct_c.h and small
// host/device helpers shared by the kernels.  Not part of the public interface.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "vct/vct_c.h"

namespace vct {

void set_error(const char* fmt, ...);

#define VCT_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t e_ = (expr);                                                                        \
    if (e_ != cudaSuccess) {                                                                        \
      ::vct::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_));  \
      return e_ == cudaErrorMemoryAllocation ? VCT_ERR_OOM : VCT_ERR_CUDA;                          \
    }                                                                                               \
  } while (0)

#define VCT_REQUIRE(cond, msg)                                    \
  do {                                                            \
    if (!(cond)) {                                                \
      ::vct::set_error("%s: %s", __func__, msg);                  \
      return VCT_ERR_INVALID;                                     \
    }                                                             \
  } while (0)

// ---- device-side records -------------------------------------------------------------

// per-draw record built on the host when the draw list is set
struct DrawRec {
  uint32_t first_index, index_count, vertex_base, material;
  float model[16];
  float nmat[9];      // mat3(transpose(inverse(model))), column-major (rule R9)
  uint32_t tri_base;  // global sequence number of the draw's first triangle
  uint32_t pad[2];
};

// a rasterisable triangle after setup (rules R1-R3), shared by the voxelizer and the camera pass
struct RasterTri {
  int32_t X[3], Y[3];   // 24.8 fixed-point window coordinates
  int32_t sign;         // orientation (+1 / -1), 0 = rejected
  int32_t imin, imax, jmin, jmax;
  int32_t mshift;       // log2 of this triangle's work-item size in pixels (raster.cuh)
  long long area;       // > 0
};

struct VoxTri {        // voxelizer per-triangle record
  RasterTri rt;
  float wp[3][3];      // world position / cube_size per vertex (voxelize.vert:26)
  float nn[3][3];      // normalised normals per vertex (voxelize.vert:28)
  uint32_t material;
  uint32_t axis;
};

struct CamTri {        // camera-pass per-triangle record
  RasterTri rt;
  float world[3][3];
  float nn[3][3];
  float iw[3];
  float zw[3];
  uint32_t material;
  uint32_t pad;
};

struct FragRec {       // one voxelization fragment (32 bytes)
  uint32_t next;       // 1-based index of the next fragment in the voxel's list, 0 = end
  uint32_t voxel;
  unsigned long long key;  // canonical order key: (triangle sequence << 24) | (row << 12) | column
  float val[4];        // colour * 255 (voxelize.frag:99)
};

struct Lights {
  vct_point_light_t l[VCT_MAX_POINT_LIGHTS];
  int32_t n;
};

// level pointers of the directional pyramid
#define VCT_MAX_LEVELS 12
struct GridView {
  uint32_t* base;                 // level 0: R^3 u32, [z][y][x]
  // Levels 1..: ONE mipmapped 3-D RGBA8 array (array level k = grid level k+1) that holds the six
  // directional volumes stacked along z, each followed by a pad of zero texels (>= 1 texel at every level = the zero border of
  // CLAMP_TO_BORDER, texture_3d.cpp:10-12): direction d occupies z' in [d/6, d/6 + tex_zs) of the normalised depth.  One array
  // means ONE texture object for every fetch -- a warp-uniform handle, so a fetch is a plain TEX instead of a per-lane handle
  // "waterfall" loop -- and the direction is chosen by the z coordinate.
  cudaTextureObject_t tex_lin;    // trilinear + mip-linear (GL_LINEAR_MIPMAP_LINEAR, texture_3d.cpp:14)
  cudaTextureObject_t tex_one;    // same texels, NEAREST mip level: one level per fetch, half the filter work when the LOD is integral
  cudaTextureObject_t tex_pt;     // same texels, unfiltered (point, nearest mip, raw bytes): the software sampler's texel fetch
  float tex_zs;                   // z' = z * tex_zs + d / 6
  int pitch[VCT_MAX_LEVELS];      // level l >= 1: direction d starts at array slice d * pitch[l]; the array level is 6 * pitch[l] deep
  const uint32_t* docc[VCT_MAX_LEVELS];  // per level: dilated occupancy bits, see occ_word_index()
  // the same bits for the production march: all levels live in ONE allocation, docc[l] == docc_all + occ_tab[l].x, and the
  // per-level constants of the lookup come with one 16-byte load: {word offset, N + 1, words per row, float bits of N}
  const uint32_t* docc_all;
  uint4 occ_tab[VCT_MAX_LEVELS];
  int R, levels;
  // RGBA16F storage variant (vct_grid_create_ex): f16[level * 6 + dir] = linear (R >> level)^3 texels of four halves (level 0: the same
  // buffer for every dir = base); nullptr for RGBA8 grids
  const unsigned long long* const* f16;
};

// Occupancy bits (built by the mip stage, read by the cone tracer to skip all-zero filter footprints):
//   occ[l]  bit (x,y,z) at flat index (z*N + y)*N + x               = texel (x,y,z) of level l is non-zero in any direction
//   docc[l] bit (x+1,y+1,z+1) of a (N+1)^3 volume, rows of occ_wpr(N) words, for x,y,z in [-1, N-1]
//                                                                   = any texel of the 2x2x2 footprint whose low corner is (x,y,z) is non-zero
__host__ __device__ __forceinline__ int occ_wpr(int N) { return (N + 32) / 32; }
__host__ __device__ __forceinline__ size_t occ_words(int N) { size_t n = (size_t)N * N * N; return (n + 31) / 32; }
__host__ __device__ __forceinline__ size_t docc_words(int N) { return (size_t)(N + 1) * (N + 1) * occ_wpr(N); }

// ---- multi-GPU exchange over NVLink peer memory (csrc/peer.cu) ------------------------------------
#define VCT_MAX_RANKS 16
enum { PEER_FLAG_PUSHED = 0, PEER_FLAG_FRAME = 1, PEER_FLAG_KINDS = 2 };
// what a kernel needs to push results into the peers and to signal them
struct PeerView {
  int rank, nranks;              // nranks <= 1: not connected (single GPU)
  uint32_t epoch;                // frame number (> 0), the value written into the flags
  uint32_t* base[VCT_MAX_RANKS];   // level-0 buffer of THIS frame on every rank (own entry = local pointer)
  uint32_t* frame[VCT_MAX_RANKS];  // RGBA8 frame on every rank
  uint32_t* flags[VCT_MAX_RANKS];  // flag block of every rank: [PEER_FLAG_KINDS][VCT_MAX_RANKS] epochs, written by the source rank
  uint32_t* done_counter;        // local: blocks of the signalling kernel that have finished
  int frame_root;                // rank that receives the finished tiles (-1: every rank)
  // sparse bookkeeping across ranks (nullptr: not in use).  touched[p] = the mip tile flags of THIS frame's level-0 buffer on rank p: the rank
  // that stores a voxel into a peer's grid also marks the peer's tile, so every rank's mip build stays sparse; pushed / pushed_n = the
  // voxels this rank stored this frame -- two frames later, when the same buffer comes round again, it zeroes exactly those on every rank
  // (and un-marks their tiles) instead of every rank clearing 4 R^3 bytes.
  uint8_t* touched[VCT_MAX_RANKS];
  uint32_t* pushed;
  uint32_t* pushed_n;
  uint32_t pushed_capacity;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Called by every thread at the end of a kernel that pushed data to the peers: the LAST block to arrive
// publishes `epoch` in slot [kind][rank] of every destination's flag block.  All peer stores of the kernel are
// ordered before the flag by the system-scope fence + release store.
__device__ __forceinline__ void peer_signal_last_block(const PeerView& pv, int kind, int only_rank) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned total = gridDim.x * gridDim.y * gridDim.z;
    if (atomicAdd(pv.done_counter + kind, 1u) == total - 1u) {
      pv.done_counter[kind] = 0u;
      __threadfence_system();
      for (int p = 0; p < pv.nranks; p++)
        if (only_rank < 0 || p == only_rank) st_release_sys(pv.flags[p] + kind * VCT_MAX_RANKS + pv.rank, pv.epoch);
    }
  }
}

// Multi-GPU split of the frame: which rank owns the 32x32 screen tile (tx, ty).  A diagonal lattice -- (tx + k ty) mod n with k coprime to n --
// spreads every rank's tiles evenly in both directions; plain round-robin over the row-major tile index gives (tx + 60 ty) mod 8 =
// (tx + 4 ty) mod 8 at 1920 pixels: two column phases only, which line up with the walls of the scene (cone kernel 117..128 us per rank at N = 8).
// k is computed once per launch on the host (screen_tile_k) and travels with the kernel arguments: one modulo per test on the device.
__host__ __device__ __forceinline__ int screen_tile_k(int nranks) { return nranks % 3 ? 3 : (nranks % 5 ? 5 : 7); }
__host__ __device__ __forceinline__ int screen_tile_owner(int tx, int ty, int nranks, int k) { return (tx + k * ty) % nranks; }

// surface handles of the stacked mipmapped array: s[grid level], level >= 1; direction d starts at z = d * pitch[level]
struct SurfSet {
  cudaSurfaceObject_t s[VCT_MAX_LEVELS];
  int pitch[VCT_MAX_LEVELS];
};
// one texel / four texels of direction d
template <class T>
__device__ __forceinline__ void surf_write(const SurfSet& surf, int d, int level, T v, int x_bytes, int y, int z) {
  surf3Dwrite(v, surf.s[level], x_bytes, y, z + d * surf.pitch[level]);
}

}  // namespace vct

// ---- the objects behind the opaque handles -----------------------------------------------

struct vct_device {
  int ordinal = 0;
  cudaStream_t stream = nullptr;
  cudaDeviceProp prop{};
  // voxelizer arenas
  vct::FragRec* frags = nullptr;
  uint8_t* fresh = nullptr;          // per arena slot: the fragment was the first of its voxel (= one mark per occupied voxel)
  uint64_t frag_capacity = 0;
  int accum_mode = 0;                        // VCT_ACCUM_ORDERED (the reference's running average, default) / VCT_ACCUM_FIXED_POINT
  unsigned long long* accum = nullptr;       // fixed-point mode: two 64-bit accumulators per arena slot (all zero between frames)
  uint64_t accum_capacity = 0;
  // raster scratch (grown on demand)
  // one set per rasteriser (0 = voxelizer, 1 = G-buffer): the two run concurrently on different streams
  struct RasterScratch {
    void* tri_recs = nullptr;  size_t tri_recs_bytes = 0;
    uint32_t* item_local = nullptr;    // per-triangle exclusive prefix inside its 256-block
    uint32_t* item_block = nullptr;    // per-block exclusive prefix
    uint32_t* big_slot = nullptr;      // camera pass: where a triangle's records are (see cam_setup_kernel)
    size_t item_capacity_tris = 0;
  } rs[2];
  // second stream: the G-buffer pass of a frame does not depend on the voxel grid and runs beside clear + voxelize + mip
  cudaStream_t stream2 = nullptr;
  cudaStream_t stream3 = nullptr;    // lowest priority: the trace of a pipelined frame (trace_low_priority)
  cudaEvent_t ev_front = nullptr, ev_trace = nullptr;
  bool trace_low_priority = false;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_g0 = nullptr, ev_g1 = nullptr;
  uint32_t* counters = nullptr;      // device counters, see enum below
  uint32_t* counters_host = nullptr; // pinned mirror
  // Status words in MAPPED pinned host memory, written by kernels only when something went wrong (no traffic otherwise) and
  // polled by the host at the top of every frame / download call without touching the stream:
  //   [STATUS_OVERFLOW] fragments the last voxelization wanted when the arena was too small   [STATUS_PEER] 1 + rank a flag wait timed out on
  volatile uint32_t* status_host = nullptr;
  uint32_t* status_dev = nullptr;
  cudaEvent_t ev[8] = {};
  vct_grid* vox_owner = nullptr;      // the grid whose occupied voxels the arena's `fresh` marks describe (last vct_voxelize)
  bool have_timings = false;
  bool gbuffer_overlapped = false;   // the last frame ran its G-buffer pass on stream2 (ev_g0..ev_g1)
  // measurement / test switches (vct_debug_set); never read from the environment
  bool debug_mip_dense = false;      // every mip build reads and writes every tile
  int debug_cone_variant = -1;       // -1 = automatic; 0 literal loop, 1 two-level fetches, 2 one warp per cone slot, 3 grouped diffuse cones
  int debug_small_limit = -1;        // >= 0: bounding-box size (pixels) up to which the set-up kernels rasterise a triangle by its own lane (default: kSmallPixels / kSmallCamPixels)
  int cone_ctas_per_sm = 0;          // > 0: the persistent cone kernel launches this many CTAs per SM instead of all that fit (frames in flight: room for the other frame's kernels)
  int cone_reserved_sms = 0;         // SMs the persistent cone kernel leaves free (frames in flight: the next frame's front half runs there)
  bool debug_cone_grid = false;      // cone kernel on a host-sized grid instead of the persistent work queue
  bool mip_attr_set = false;
  // multi-GPU connection (vct_peer_connect)
  vct::PeerView peers{};               // peers.nranks <= 1 when not connected
  size_t peer_flag_bytes = 0;
  uint32_t* peer_flags = nullptr;      // local flag block [PEER_FLAG_KINDS][VCT_MAX_RANKS] + done counters + error word
  void* peer_mapped[3 * VCT_MAX_RANKS + VCT_MAX_RANKS] = {};  // pointers opened with cudaIpcOpenMemHandle (to close)
  int n_peer_mapped = 0;
  uint32_t* peer_base_all[2][VCT_MAX_RANKS] = {};   // both level-0 buffers of every rank
  uint8_t* peer_touched_all[2][VCT_MAX_RANKS] = {}; // the mip tile flags of both level-0 buffers on every rank (nullptr: dense clear + dense mip)
  uint32_t* pushed_list[2] = {};                    // voxels this rank stored into buffer 0 / 1 the last time it was the frame's buffer
  uint32_t* pushed_n = nullptr;                     // [4]: ring of counts indexed by frame number & 3 (peer.cu)
  size_t pushed_capacity = 0;
  int peer_sparse_mode = -1;                        // -1 undecided, 0 dense clear + dense mip under peers, 1 sparse (decided at the first frame of a connection)
  vct_grid* peer_grid = nullptr;
  vct_target_t_* peer_target = nullptr;
  int peer_replicate = -1;             // -1 undecided, 1 = small scene: every rank voxelizes all of it (no voxel exchange), 0 = z-slabs + push
  int peer_replicate_force = -1;       // VCT_DEBUG_PEER_REPLICATE: -1 automatic, 0 never, 1 always
  uint32_t peer_epoch = 0;             // frames rendered since vct_peer_connect
  bool peer_export_fresh = false;      // vct_peer_export ran (flag block zeroed) and no connect has consumed it yet
};
enum { STATUS_OVERFLOW = 0, STATUS_PEER = 1, STATUS_WORDS = 4 };
enum { CNT_ITEMS = 0, CNT_FRAGS = 1, CNT_OCCUPIED = 2, CNT_MAXLIST = 3, CNT_CAM_RECS = 11 /* records of the camera pass */, CNT_CAM_ITEMS = 12 /* outside the words the voxelizer clears every frame */, CNT_TICKET_VOX = 13, CNT_TICKET_CAM = 14 /* last-block tickets of the setup kernels */, CNT_CONE_WORK = 15 /* next item of the persistent cone kernel */, CNT_SAMPLES = 16 /* ..31, as 8 x u64 */, CNT_TOTAL = 32 };

struct vct_scene {
  vct_device* dev = nullptr;
  // the arrays the kernels read (= the current halves of the double buffers below)
  vct_vertex_t* verts = nullptr;  uint32_t n_verts = 0;
  uint32_t* indices = nullptr;    uint32_t n_indices = 0;
  vct_material_t* mats = nullptr; uint32_t n_mats = 0;
  vct::DrawRec* draws = nullptr;  uint32_t n_draws = 0;
  // Every array is double buffered and uploaded on its own stream: vct_scene_set_* writes the half that no queued frame reads
  // (it waits for the event recorded when that half was retired) and flips, so the host-to-device copies of frame i+1 run while
  // frame i renders instead of at the head of frame i+1 (54 us per frame in bench.py's end-to-end loop).  The frame stream
  // waits for `uploaded` before its next use of the scene (scene_ready()).
  struct DBuf { void* buf[2] = {}; size_t cap_bytes[2] = {}; int cur = 0; cudaEvent_t retired[2] = {}; };
  enum { DB_VERTS = 0, DB_INDICES = 1, DB_MATS = 2, DB_DRAWS = 3, DB_COUNT = 4 };
  DBuf db[DB_COUNT];
  cudaStream_t upload_stream = nullptr;
  cudaEvent_t uploaded = nullptr;
  bool upload_pending = false;
  uint32_t n_tris = 0;
  vct::Lights lights{};
  float cube_size = 1.0f;
  // pinned staging so that per-frame updates are true async copies: a ring of bump-allocated slots, each guarded by the
  // event of its last copy, so that an upload never waits for the frames still in flight on the stream
  struct StageSlot { unsigned char* buf = nullptr; size_t cap = 0, used = 0; cudaEvent_t done = nullptr; bool pending = false; };
  static constexpr int kStageSlots = 3;
  StageSlot stage[kStageSlots];
  int stage_cur = 0;
};

struct vct_grid {
  vct_device* dev = nullptr;
  int R = 0, levels = 0;
  int fmt = 0;                          // VCT_GRID_RGBA8 / VCT_GRID_RGBA16F
  unsigned long long* f16_lvl[VCT_MAX_LEVELS][6] = {};   // RGBA16F: levels 1.. per direction (linear); level 0 = base (8 bytes per voxel)
  const unsigned long long** f16_table = nullptr;        // the same pointers in device memory: [level * 6 + dir]
  uint32_t* base = nullptr;             // level 0 of the current frame
  uint32_t* base_buf[2] = {};           // base_buf[0] == the allocation; base_buf[1] only in multi-GPU mode (double buffering)
  size_t bytes = 0;
  // levels 1..: ONE mipmapped array, the six directions stacked along z (see GridView); written by the mip kernels through surfaces
  cudaMipmappedArray_t marr = nullptr;
  cudaTextureObject_t tex_lin = 0, tex_one = 0, tex_pt = 0;
  float tex_zs = 0.0f;
  vct::SurfSet surf{};
  // ---- scratch of the mip kernels (csrc/mipmap.cu): small linear copies of the coarse levels that cross warps / CTAs ----
  uint32_t* rec3 = nullptr;             // level 3 as records of six words per texel: what a 32^3 block's last tile needs of the other tiles
  uint32_t* rec_top = nullptr;          // levels 5.. as records (level l at rec_top + top_off[l]): input of the single-CTA top of the chain
  uint32_t top_off[VCT_MAX_LEVELS] = {};
  uint8_t* occb = nullptr;              // levels 3..: one occupancy byte per texel (level l at occb + occb_off[l])
  uint32_t occb_off[VCT_MAX_LEVELS] = {};
  uint32_t* sb_epoch = nullptr;         // per 64^3 super-block: the last mip build (mip_build) that processed one of its tiles
  uint32_t mip_build = 0;
  uint32_t* mip_counters = nullptr;     // [0] = blocks finished, [1 + b] = tiles of 32^3 block b that have arrived
  uint8_t* tile_zero = nullptr;         // mip stage: per 32x8x8 tile "every output of this tile is known to be zero" (skip rewriting zeros)
  // ---- sparse frame-to-frame bookkeeping (< 1 % of the voxels are occupied) ----
  // tile_touched[tile] != 0: the last vct_voxelize wrote a voxel of this 32x8x8 tile.  While flags_valid, every non-zero word of
  // level 0 lies in a touched tile, so the mip build skips untouched tiles whose outputs are already zero without reading them.
  // While sparse_clear_ok (and dev->vox_owner == this), the non-zero words are exactly the device's occupied list and
  // vct_grid_clear zeroes those instead of the whole level.  Anything that writes level 0 behind the library's back
  // (vct_grid_upload_base, the raw pointer) drops back to the dense paths.
  uint8_t* tile_touched = nullptr;
  const uint8_t* peer_touched = nullptr;   // multi-GPU frame: the tile flags of this frame's level-0 buffer, kept by the pushing ranks (peer.cu)
  bool flags_valid = false, sparse_clear_ok = false, base_zero = false, external = false;
  int dirty_z0 = 0, dirty_z1 = 0;       // z range voxelized (or uploaded) since the last clear: vct_voxelize refuses a slab that overlaps it
  void untrack() { flags_valid = sparse_clear_ok = base_zero = false; dirty_z0 = 0; dirty_z1 = R; }
  uint32_t* occ[VCT_MAX_LEVELS] = {};   // non-dilated occupancy bits per level (levels 4.. point into occ_hi)
  uint32_t* occ_hi = nullptr;
  uint32_t occ_hi_words = 0, docc_hi_off = 0, docc_hi_words = 0;   // what the fused mip kernel zeroes every build: plain / dilated bits of levels 4..
  uint32_t* docc[VCT_MAX_LEVELS] = {};  // dilated occupancy bits per level: pointers into docc_all
  uint32_t* docc_all = nullptr;
  uint32_t docc_off[VCT_MAX_LEVELS] = {};   // word offset of each level in docc_all
  vct::GridView view() const {
    vct::GridView v;
    v.base = base; v.R = R; v.levels = levels;
    for (int i = 0; i < VCT_MAX_LEVELS; i++) { v.docc[i] = docc[i]; v.pitch[i] = surf.pitch[i]; }
    v.tex_lin = tex_lin; v.tex_one = tex_one; v.tex_pt = tex_pt; v.tex_zs = tex_zs;
    v.docc_all = docc_all;
    v.f16 = f16_table;
    for (int l = 0; l < VCT_MAX_LEVELS; l++) {
      const int N = l < levels ? (R >> l) : 0;
      const float fN = (float)N;
      uint32_t bits;
      memcpy(&bits, &fN, 4);
      v.occ_tab[l] = make_uint4(docc_off[l], (uint32_t)(N + 1), (uint32_t)vct::occ_wpr(N), bits);
    }
    return v;
  }
};

struct vct_target_t_ {
  vct_device* dev = nullptr;
  int W = 0, H = 0;
  unsigned long long* vis = nullptr;  // (depth bits << 32) | triangle sequence, atomicMin
  float* world_pos = nullptr;         // 3 floats / pixel
  float* normal = nullptr;            // 3 floats / pixel (interpolated, NOT renormalised)
  uint32_t* material = nullptr;
  uint32_t* frame = nullptr;          // RGBA8
  void* cone_out = nullptr;           // float4 [slot][pixel], grown on demand by the cone tracer
  size_t cone_out_elems = 0;
  uint32_t* tile_list = nullptr;      // [0] = count, [1..] = live 8x4 tiles
  // asynchronous read-back (vct_target_download_frame_async): device-side snapshots + a copy stream
  uint32_t* snap[2] = {};
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t snap_ready[2] = {}, copy_done[2] = {};
  uint64_t n_async = 0;               // tickets issued
  // a read-back whose device-to-host copy has not been enqueued yet: vct_render_frame starts it together with its cone kernel
  uint32_t* pending_host = nullptr;
  int pending_slot = 0;
  cudaEvent_t copy_gate = nullptr;    // "the next frame's cone kernel is about to start"
};

struct vct_tex3d {
  vct_device* dev = nullptr;
  int w = 0, h = 0, d = 0, levels = 0;
  uint32_t* lvl[VCT_MAX_LEVELS] = {};
};

static inline unsigned grid_for(size_t n, unsigned threads = 256, unsigned cap = 148 * 16) {
  size_t b = (n + threads - 1) / threads;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ---- stage entry points implemented in the .cu files ------------------------------------
namespace vct {
int launch_sparse_clear(vct_device* dev, vct_grid* g);
int scene_ready(vct_scene* sc);
int launch_copy_u32(cudaStream_t s, uint32_t* dst, const uint32_t* src, size_t n);
int ensure_tri_scratch(vct_device* dev, int which /* 0 voxelizer, 1 G-buffer */, size_t n_tris, size_t rec_bytes_total);
int launch_voxelize(vct_device* dev, vct_scene* sc, vct_grid* g, int z0, int z1, const PeerView* push = nullptr);
int launch_peer_wait(vct_device* dev, int kind, uint32_t epoch);
int launch_peer_signal(vct_device* dev, const PeerView& pv, int kind);      // publish pv.epoch under `kind` to every rank (a launch of its own)
int launch_peer_unpush(vct_device* dev, const PeerView& pv, int logR);   // zero the voxels this rank pushed into this frame's buffer two frames ago
int ensure_pushed_lists(vct_device* dev);
int check_status(vct_device* dev);   // VCT_ERR_OVERFLOW / VCT_ERR_CUDA if a kernel reported an arena overflow / a peer timeout since the last check
int launch_mipmap(vct_device* dev, vct_grid* g);
bool mip_fused_applies(int R, int levels);   // the fused mip kernel (32x8x8 tiles, 32^3 blocks) handles this grid
// tile_list: the resolve kernel also builds the cone tracer's live-tile list (and resets its work counter): launch_cone_trace phase 2 may follow directly
int launch_gbuffer(vct_device* dev, vct_scene* sc, const float* view, const float* proj, vct_target_t_* t, int tile_rank = 0, int tile_nranks = 1, bool tile_list = false);
// phase: 0 = tile list + cones + shade; 1 = the live-tile list only (depends on the G-buffer alone: vct_render_frame builds it on the
// G-buffer stream); 2 = cones + shade with the list of a preceding phase-1 call
int launch_cone_trace(vct_device* dev, vct_scene* sc, vct_grid* g, const float* view, const vct_trace_params_t* p, vct_target_t_* t,
                      bool count_samples, const vct::PeerView* push = nullptr, int phase = 0);
int launch_fill_u32(cudaStream_t s, uint32_t* p, size_t n, uint32_t v);
int launch_tex3d_mip(vct_tex3d* t);
}  // namespace vct
