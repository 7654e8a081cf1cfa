// capi.cu -- the extern "C" layer of include/vct/vct_c.h: object lifetime, uploads/downloads and
// the stream-ordered frame sequence.  No compute lives here; there is no CPU fallback anywhere.
#include <stdarg.h>

#include <new>

#include "vct_internal.cuh"

namespace vct {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

// R9: mat3(transpose(inverse(model))) for an affine model matrix = cofactor / det, in double
static void normal_matrix(const float* m, float* nm) {
  double a00 = m[0], a10 = m[1], a20 = m[2];
  double a01 = m[4], a11 = m[5], a21 = m[6];
  double a02 = m[8], a12 = m[9], a22 = m[10];
  volatile double c00 = a11 * a22 - a12 * a21, c01 = a12 * a20 - a10 * a22, c02 = a10 * a21 - a11 * a20;
  volatile double c10 = a02 * a21 - a01 * a22, c11 = a00 * a22 - a02 * a20, c12 = a01 * a20 - a00 * a21;
  volatile double c20 = a01 * a12 - a02 * a11, c21 = a02 * a10 - a00 * a12, c22 = a00 * a11 - a01 * a10;
  volatile double d0 = a00 * c00, d1 = a01 * c01, d2 = a02 * c02;
  double det = (d0 + d1) + d2;
  nm[0] = (float)(c00 / det); nm[1] = (float)(c10 / det); nm[2] = (float)(c20 / det);
  nm[3] = (float)(c01 / det); nm[4] = (float)(c11 / det); nm[5] = (float)(c21 / det);
  nm[6] = (float)(c02 / det); nm[7] = (float)(c12 / det); nm[8] = (float)(c22 / det);
}

template <class T>
static int ensure_capacity(T** p, size_t* cap, size_t n) {
  if (n <= *cap) return VCT_OK;
  if (*p) cudaFree(*p);
  *p = nullptr; *cap = 0;
  size_t want = n + n / 4 + 64;
  VCT_CUDA(cudaMalloc((void**)p, want * sizeof(T)));
  *cap = want;
  return VCT_OK;
}

// Host -> device copy of `bytes` through pinned staging, asynchronous on the device stream.  The staging memory is a ring of
// slots filled front to back; a slot is reused only after the event recorded behind its last copy has completed (normally
// long ago), so per-frame scene updates do not serialise the host with the frames queued on the device.
static int stage_upload(vct_scene* sc, void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return VCT_OK;
  cudaStream_t s = sc->upload_stream;
  vct_scene::StageSlot* slot = &sc->stage[sc->stage_cur];
  const size_t need = (bytes + 255) & ~(size_t)255;
  if (slot->used + need > slot->cap) {
    sc->stage_cur = (sc->stage_cur + 1) % vct_scene::kStageSlots;
    slot = &sc->stage[sc->stage_cur];
    if (slot->pending) { VCT_CUDA(cudaEventSynchronize(slot->done)); slot->pending = false; }
    slot->used = 0;
    if (slot->cap < need) {
      if (slot->buf) cudaFreeHost(slot->buf);
      slot->buf = nullptr; slot->cap = 0;
      const size_t want = need * 2 > ((size_t)1 << 20) ? need * 2 : ((size_t)1 << 20);
      VCT_CUDA(cudaMallocHost((void**)&slot->buf, want));
      slot->cap = want;
    }
    if (!slot->done) VCT_CUDA(cudaEventCreateWithFlags(&slot->done, cudaEventDisableTiming));
  }
  memcpy(slot->buf + slot->used, src, bytes);
  VCT_CUDA(cudaMemcpyAsync(dst, slot->buf + slot->used, bytes, cudaMemcpyHostToDevice, s));
  VCT_CUDA(cudaEventRecord(slot->done, s));
  slot->pending = true;
  slot->used += need;
  return VCT_OK;
}

// Uploads `bytes` into the idle half of double buffer `which` on the upload stream and makes it the current half.
static int upload_array(vct_scene* sc, int which, const void* src, size_t bytes, void** current) {
  vct_scene::DBuf& d = sc->db[which];
  const int nb = d.cur ^ 1;
  if (bytes > d.cap_bytes[nb]) {
    if (d.buf[nb]) cudaFree(d.buf[nb]);   // (cudaFree waits for the device: no queued frame can still read it)
    d.buf[nb] = nullptr; d.cap_bytes[nb] = 0;
    const size_t want = bytes + bytes / 4 + 256;
    VCT_CUDA(cudaMalloc(&d.buf[nb], want));
    d.cap_bytes[nb] = want;
  }
  // the idle half was read by the frames queued before it was retired: the copy waits for them, not for the frames queued since
  if (d.retired[nb]) VCT_CUDA(cudaStreamWaitEvent(sc->upload_stream, d.retired[nb], 0));
  int rc = stage_upload(sc, d.buf[nb], src, bytes);
  if (rc) return rc;
  if (!d.retired[d.cur]) VCT_CUDA(cudaEventCreateWithFlags(&d.retired[d.cur], cudaEventDisableTiming));
  VCT_CUDA(cudaEventRecord(d.retired[d.cur], sc->dev->stream));   // everything queued so far may read the half being retired
  d.cur = nb;
  *current = d.buf[nb];
  VCT_CUDA(cudaEventRecord(sc->uploaded, sc->upload_stream));
  sc->upload_pending = true;
  return VCT_OK;
}

// called by every entry point that launches kernels reading the scene: the frame stream waits for the uploads in flight
int scene_ready(vct_scene* sc) {
  if (!sc->upload_pending) return VCT_OK;
  VCT_CUDA(cudaStreamWaitEvent(sc->dev->stream, sc->uploaded, 0));
  sc->upload_pending = false;
  return VCT_OK;
}

// What the kernels reported through the mapped status words since the last call (host-side poll, nothing is enqueued):
//  * arena overflow: the voxelization of an earlier frame dropped fragments (which ones depends on atomic order).  The arena is
//    grown to what that frame wanted (+25 %) and VCT_ERR_OVERFLOW is returned ONCE: the caller re-issues the frame
//    (vct::Renderer::render does).  The frame that overflowed is reported at the next call that sees the word, i.e. one or two
//    frames later when frames are queued asynchronously, immediately after a vct_device_sync / blocking download.
//  * peer timeout: a flag wait of the multi-GPU exchange gave up (csrc/peer.cu): the frame was built from incomplete data.
int check_status(vct_device* dev) {
  if (!dev->status_host) return VCT_OK;
  if (const uint32_t r = dev->status_host[STATUS_PEER]) {
    dev->status_host[STATUS_PEER] = 0u;
    set_error("peer wait timed out: rank %u never signalled (peer process dead, not connected, or frames out of step); the last frame is incomplete", r - 1u);
    return VCT_ERR_CUDA;
  }
  if (const uint32_t wanted = dev->status_host[STATUS_OVERFLOW]) {
    dev->status_host[STATUS_OVERFLOW] = 0u;
    const uint64_t cap = dev->frag_capacity, want = (uint64_t)wanted + wanted / 4;
    set_error("fragment arena overflow: a voxelization produced %u fragments, capacity %llu; the arena has been grown, re-issue the frame", wanted,
              (unsigned long long)cap);
    if (want > cap) {
      int rc = vct_voxelize_reserve(dev, want < 0xFFFFFFF0ull ? want : 0xFFFFFFEFull);
      if (rc) return rc;
    }
    return VCT_ERR_OVERFLOW;
  }
  return VCT_OK;
}

}  // namespace vct

using namespace vct;

extern "C" {

const char* vct_last_error(void) { return g_err; }
const char* vct_version(void) { return "vct-b200 0.1 (sm_100a)"; }

// ------------------------------------------------------------------ device
int vct_device_create(int ordinal, vct_device_t** out) {
  VCT_REQUIRE(out != nullptr, "out is null");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error("no CUDA device available (%s); this library has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "count = 0");
    return VCT_ERR_CUDA;
  }
  VCT_REQUIRE(ordinal >= 0 && ordinal < n, "bad device ordinal");
  VCT_CUDA(cudaSetDevice(ordinal));
  vct_device* d = new (std::nothrow) vct_device();
  if (!d) { set_error("out of host memory"); return VCT_ERR_OOM; }
  d->ordinal = ordinal;
  VCT_CUDA(cudaGetDeviceProperties(&d->prop, ordinal));
  if (d->prop.major < 10) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a only", ordinal, d->prop.major, d->prop.minor);
    delete d;
    return VCT_ERR_CUDA;
  }
  // the frame's critical path (clear -> voxelize -> mip -> trace) runs on a high-priority stream, the G-buffer pass beside it on a
  // low-priority one: when both have blocks ready the scheduler serves the critical path first
  int prio_lo = 0, prio_hi = 0;
  VCT_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  VCT_CUDA(cudaStreamCreateWithPriority(&d->stream, cudaStreamNonBlocking, prio_hi));
  VCT_CUDA(cudaMalloc(&d->counters, CNT_TOTAL * sizeof(uint32_t)));
  VCT_CUDA(cudaMemset(d->counters, 0, CNT_TOTAL * sizeof(uint32_t)));
  VCT_CUDA(cudaMallocHost(&d->counters_host, CNT_TOTAL * sizeof(uint32_t)));
  VCT_CUDA(cudaHostAlloc((void**)&d->status_host, STATUS_WORDS * sizeof(uint32_t), cudaHostAllocMapped));
  for (int i = 0; i < STATUS_WORDS; i++) d->status_host[i] = 0u;
  VCT_CUDA(cudaHostGetDevicePointer((void**)&d->status_dev, (void*)d->status_host, 0));
  for (int i = 0; i < 8; i++) VCT_CUDA(cudaEventCreate(&d->ev[i]));
  // numerically lower = more urgent: [prio_hi, prio_lo].  The G-buffer stream sits between the critical path and the trace stream
  VCT_CUDA(cudaStreamCreateWithPriority(&d->stream2, cudaStreamNonBlocking, prio_hi < prio_lo ? prio_hi + 1 : prio_lo));
  // one trace stream per process and device ordinal: the traces of all pipelines on a GPU run one after the other (a persistent
  // cone kernel launched beside another one would find only the reserved SMs free and retire without doing its work)
  static cudaStream_t g_trace_stream[64] = {};
  if (!g_trace_stream[ordinal & 63]) VCT_CUDA(cudaStreamCreateWithPriority(&g_trace_stream[ordinal & 63], cudaStreamNonBlocking, prio_lo));
  d->stream3 = g_trace_stream[ordinal & 63];
  VCT_CUDA(cudaEventCreateWithFlags(&d->ev_front, cudaEventDisableTiming));
  VCT_CUDA(cudaEventCreateWithFlags(&d->ev_trace, cudaEventDisableTiming));
  VCT_CUDA(cudaEventCreateWithFlags(&d->ev_fork, cudaEventDisableTiming));
  VCT_CUDA(cudaEventCreateWithFlags(&d->ev_join, cudaEventDisableTiming));
  VCT_CUDA(cudaEventCreate(&d->ev_g0));
  VCT_CUDA(cudaEventCreate(&d->ev_g1));
  *out = d;
  return VCT_OK;
}

int vct_device_destroy(vct_device_t* d) {
  if (!d) return VCT_OK;
  cudaSetDevice(d->ordinal);
  vct_peer_disconnect(d);
  cudaStreamSynchronize(d->stream);
  cudaFree(d->peer_flags);
  cudaFree(d->frags); cudaFree(d->fresh); cudaFree(d->accum);
  for (auto& r : d->rs) { cudaFree(r.tri_recs); cudaFree(r.item_local); cudaFree(r.item_block); cudaFree(r.big_slot); }
  if (d->stream2) { cudaStreamSynchronize(d->stream2); cudaStreamDestroy(d->stream2); }
  if (d->stream3) cudaStreamSynchronize(d->stream3);   // shared by the device objects of this GPU: never destroyed
  for (cudaEvent_t e : {d->ev_fork, d->ev_join, d->ev_g0, d->ev_g1, d->ev_front, d->ev_trace}) if (e) cudaEventDestroy(e);
  cudaFree(d->counters); cudaFreeHost(d->counters_host); cudaFreeHost((void*)d->status_host);
  for (int i = 0; i < 8; i++) if (d->ev[i]) cudaEventDestroy(d->ev[i]);
  cudaStreamDestroy(d->stream);
  delete d;
  return VCT_OK;
}

int vct_device_sync(vct_device_t* d) {
  VCT_REQUIRE(d, "device is null");
  VCT_CUDA(cudaStreamSynchronize(d->stream));
  return check_status(d);
}

void* vct_device_stream(vct_device_t* d) { return d ? (void*)d->stream : nullptr; }

// ------------------------------------------------------------------ scene
int vct_scene_create(vct_device_t* dev, vct_scene_t** out) {
  VCT_REQUIRE(dev && out, "null argument");
  vct_scene* s = new (std::nothrow) vct_scene();
  if (!s) { set_error("out of host memory"); return VCT_ERR_OOM; }
  s->dev = dev;
  if (cudaStreamCreateWithFlags(&s->upload_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&s->uploaded, cudaEventDisableTiming) != cudaSuccess) {
    set_error("scene: stream/event creation failed");
    delete s;
    return VCT_ERR_CUDA;
  }
  *out = s;
  return VCT_OK;
}

int vct_scene_destroy(vct_scene_t* s) {
  if (!s) return VCT_OK;
  cudaStreamSynchronize(s->upload_stream);
  cudaStreamSynchronize(s->dev->stream);
  cudaStreamSynchronize(s->dev->stream2);
  for (auto& d : s->db)
    for (int i = 0; i < 2; i++) { cudaFree(d.buf[i]); if (d.retired[i]) cudaEventDestroy(d.retired[i]); }
  cudaStreamDestroy(s->upload_stream);
  cudaEventDestroy(s->uploaded);
  for (auto& slot : s->stage) {
    if (slot.buf) cudaFreeHost(slot.buf);
    if (slot.done) cudaEventDestroy(slot.done);
  }
  delete s;
  return VCT_OK;
}

int vct_scene_set_geometry(vct_scene_t* s, const vct_vertex_t* verts, uint32_t n_verts, const uint32_t* indices, uint32_t n_indices) {
  VCT_REQUIRE(s, "scene is null");
  VCT_REQUIRE((verts || !n_verts) && (indices || !n_indices), "null geometry");
  int rc;
  if ((rc = upload_array(s, vct_scene::DB_VERTS, verts, (size_t)n_verts * sizeof(vct_vertex_t), (void**)&s->verts))) return rc;
  if ((rc = upload_array(s, vct_scene::DB_INDICES, indices, (size_t)n_indices * sizeof(uint32_t), (void**)&s->indices))) return rc;
  s->n_verts = n_verts; s->n_indices = n_indices;
  return VCT_OK;
}

int vct_scene_set_materials(vct_scene_t* s, const vct_material_t* mats, uint32_t n) {
  VCT_REQUIRE(s && (mats || !n), "null argument");
  int rc;
  if ((rc = upload_array(s, vct_scene::DB_MATS, mats, (size_t)n * sizeof(vct_material_t), (void**)&s->mats))) return rc;
  s->n_mats = n;
  return VCT_OK;
}

int vct_scene_set_draws(vct_scene_t* s, const vct_draw_t* draws, uint32_t n) {
  VCT_REQUIRE(s && (draws || !n), "null argument");
  std::vector<DrawRec> recs(n);
  uint32_t tri = 0;
  for (uint32_t i = 0; i < n; i++) {
    const vct_draw_t& d = draws[i];
    VCT_REQUIRE((uint64_t)d.first_index + d.index_count <= s->n_indices, "draw range outside the index buffer");
    VCT_REQUIRE(d.material < s->n_mats, "draw references an unknown material");
    DrawRec& r = recs[i];
    memset(&r, 0, sizeof r);
    r.first_index = d.first_index; r.index_count = d.index_count; r.vertex_base = d.vertex_base; r.material = d.material;
    memcpy(r.model, d.model, sizeof r.model);
    normal_matrix(d.model, r.nmat);
    r.tri_base = tri;
    tri += d.index_count / 3;
  }
  int rc;
  if ((rc = upload_array(s, vct_scene::DB_DRAWS, recs.data(), (size_t)n * sizeof(DrawRec), (void**)&s->draws))) return rc;
  s->n_draws = n;
  s->n_tris = tri;
  return VCT_OK;
}

int vct_scene_set_lights(vct_scene_t* s, const vct_point_light_t* lights, uint32_t n) {
  VCT_REQUIRE(s && (lights || !n), "null argument");
  uint32_t m = n < VCT_MAX_POINT_LIGHTS ? n : VCT_MAX_POINT_LIGHTS;  // min(point_light_count, MAX_POINT_LIGHTS), voxelize.frag:128
  memset(&s->lights, 0, sizeof s->lights);
  for (uint32_t i = 0; i < m; i++) s->lights.l[i] = lights[i];
  s->lights.n = (int32_t)m;
  return VCT_OK;
}

int vct_scene_set_cube_size(vct_scene_t* s, float cube_size) {
  VCT_REQUIRE(s, "scene is null");
  s->cube_size = cube_size;
  return VCT_OK;
}

// ------------------------------------------------------------------ grid
int vct_grid_create(vct_device_t* dev, int R, int levels, vct_grid_t** out) { return vct_grid_create_ex(dev, R, levels, VCT_GRID_RGBA8, out); }

int vct_grid_create_ex(vct_device_t* dev, int R, int levels, int format, vct_grid_t** out) {
  VCT_REQUIRE(dev && out, "null argument");
  VCT_REQUIRE(format == VCT_GRID_RGBA8 || format == VCT_GRID_RGBA16F, "format must be VCT_GRID_RGBA8 or VCT_GRID_RGBA16F");
  const bool f16 = format == VCT_GRID_RGBA16F;
  // 1024 = the largest size the 32-bit voxel index (fragment records, tile flags, sparse clear) addresses and the tests cover
  VCT_REQUIRE(R >= 2 && R <= 1024 && (R & (R - 1)) == 0, "resolution must be a power of two in [2, 1024]");
  VCT_REQUIRE(levels >= 1 && levels <= VCT_MAX_LEVELS && (R >> (levels - 1)) >= 1, "levels must satisfy 1 <= levels <= log2(R)+1");
  vct_grid* g = new (std::nothrow) vct_grid();
  if (!g) { set_error("out of host memory"); return VCT_ERR_OOM; }
  g->dev = dev; g->R = R; g->levels = levels; g->fmt = format;
  size_t n0 = (size_t)R * R * R;
  cudaError_t e = cudaMalloc(&g->base, n0 * (f16 ? 8 : 4));
  g->base_buf[0] = g->base;
  g->bytes = n0 * (f16 ? 8 : 4);
  if (f16 && e == cudaSuccess) {
    // RGBA16F variant: levels 1.. as linear buffers per direction (no texture array: the tracer filters in fp32), pointer table on the device
    std::vector<const unsigned long long*> table((size_t)levels * 6, nullptr);
    for (int d = 0; d < 6; d++) table[d] = reinterpret_cast<const unsigned long long*>(g->base);
    for (int l = 1; l < levels && e == cudaSuccess; l++) {
      const size_t n = (size_t)(R >> l) * (R >> l) * (R >> l);
      for (int d = 0; d < 6 && e == cudaSuccess; d++) {
        e = cudaMalloc(&g->f16_lvl[l][d], n * 8);
        if (e == cudaSuccess) e = cudaMemsetAsync(g->f16_lvl[l][d], 0, n * 8, dev->stream);
        g->bytes += n * 8;
        table[(size_t)l * 6 + d] = g->f16_lvl[l][d];
      }
    }
    if (e == cudaSuccess) e = cudaMalloc(&g->f16_table, table.size() * sizeof(void*));
    if (e == cudaSuccess) e = cudaMemcpy(g->f16_table, table.data(), table.size() * sizeof(void*), cudaMemcpyHostToDevice);
  }
  if (!f16 && mip_fused_applies(R, levels)) {
    // the streaming mip kernel (csrc/mipmap.cu): one flag pair per 32x8x8 warp-tile, and the small linear copies of the coarse
    // levels that the tail kernel folds
    const size_t n_tiles = n0 / 2048, n_blocks = n0 / 32768, n3 = n0 / 512;
    auto zalloc = [&](void** p, size_t bytes) {
      if (e == cudaSuccess) e = cudaMalloc(p, bytes);
      if (e == cudaSuccess) e = cudaMemsetAsync(*p, 0, bytes, dev->stream);
      g->bytes += bytes;
    };
    zalloc((void**)&g->tile_touched, n_tiles);
    zalloc((void**)&g->tile_zero, n_tiles);           // 0 = unknown: the first build writes everything
    zalloc((void**)&g->mip_counters, (1 + n_blocks) * 4);
    zalloc((void**)&g->sb_epoch, (n0 / 262144 + 1) * 4);
    zalloc((void**)&g->rec3, n3 * 24);
    size_t top_words = 0, occ_bytes = 0;
    for (int l = 3; l < levels; l++) {
      const size_t n = (size_t)(R >> l) * (R >> l) * (R >> l);
      g->occb_off[l] = (uint32_t)occ_bytes;
      occ_bytes += (n + 15) & ~(size_t)15;
      if (l >= 5) { g->top_off[l] = (uint32_t)top_words; top_words += n * 6; }
    }
    zalloc((void**)&g->rec_top, top_words * 4);
    zalloc((void**)&g->occb, occ_bytes);
  }
  size_t docc_total = 0;   // the dilated bits of all levels share one allocation (256-byte aligned parts)
  for (int l = 0; l < levels; l++) {
    g->docc_off[l] = (uint32_t)docc_total;
    docc_total += (docc_words(R >> l) + 63) & ~(size_t)63;
  }
  if (e == cudaSuccess) e = cudaMalloc(&g->docc_all, docc_total * 4);
  if (e == cudaSuccess) e = cudaMemsetAsync(g->docc_all, 0, docc_total * 4, dev->stream);
  g->bytes += docc_total * 4;
  // plain occupancy bits: levels 0-3 on their own, levels 4.. in one allocation (the fused mip kernel zeroes that range, and the dilated
  // bits of the same levels, every build: the tail kernel ORs the occupied texels in)
  size_t occ_hi_total = 0;
  for (int l = 4; l < levels; l++) occ_hi_total += (occ_words(R >> l) + 3) & ~(size_t)3;
  if (occ_hi_total && e == cudaSuccess) {
    e = cudaMalloc(&g->occ_hi, occ_hi_total * 4);
    if (e == cudaSuccess) e = cudaMemsetAsync(g->occ_hi, 0, occ_hi_total * 4, dev->stream);
    g->bytes += occ_hi_total * 4;
  }
  g->occ_hi_words = (uint32_t)occ_hi_total;
  g->docc_hi_off = levels > 4 ? g->docc_off[4] : (uint32_t)docc_total;
  g->docc_hi_words = (uint32_t)(docc_total - g->docc_hi_off);
  size_t hi_off = 0;
  for (int l = 0; l < levels && e == cudaSuccess; l++) {
    const int N = R >> l;
    g->docc[l] = g->docc_all + g->docc_off[l];
    if (l >= 4) {
      g->occ[l] = g->occ_hi + hi_off;
      hi_off += (occ_words(N) + 3) & ~(size_t)3;
      continue;
    }
    e = cudaMalloc(&g->occ[l], occ_words(N) * 4);
    if (e == cudaSuccess) e = cudaMemsetAsync(g->occ[l], 0, occ_words(N) * 4, dev->stream);
    g->bytes += occ_words(N) * 4;
  }
  // levels 1.. additionally live in ONE mipmapped CUDA array (six directions stacked along z, zero pads between them) so that
  // the cone tracer can use the texture units with a single, warp-uniform texture object (GridView)
  if (levels >= 2 && e == cudaSuccess && !f16) {
    const int L = levels - 1;                                   // array levels
    const int n_coarse = R >> (levels - 1);                     // size of the coarsest level
    const int pitch_coarse = n_coarse + 1;                      // volume + one zero texel
    const size_t depth0 = (size_t)6 * pitch_coarse << (L - 1);  // divisible by 2^(L-1): every level's depth is exactly 6 * pitch
    cudaChannelFormatDesc fmt = cudaCreateChannelDesc<uchar4>();
    cudaExtent ext = make_cudaExtent((size_t)R / 2, (size_t)R / 2, depth0);
    e = cudaMallocMipmappedArray(&g->marr, &fmt, ext, (unsigned)L, cudaArraySurfaceLoadStore);
    for (int l = 1; l < levels && e == cudaSuccess; l++) {
      cudaArray_t arr;
      e = cudaGetMipmappedArrayLevel(&arr, g->marr, (unsigned)(l - 1));
      if (e != cudaSuccess) break;
      cudaResourceDesc rd;
      memset(&rd, 0, sizeof rd);
      rd.resType = cudaResourceTypeArray;
      rd.res.array.array = arr;
      e = cudaCreateSurfaceObject(&g->surf.s[l], &rd);
      g->surf.pitch[l] = pitch_coarse << (levels - 1 - l);
      const size_t n = (size_t)(R >> l);
      g->bytes += n * n * (size_t)6 * g->surf.pitch[l] * 4;
      // the pads must read as zero for ever (the volumes are rewritten by every mip build): zero the level in z-chunks
      if (e == cudaSuccess) {
        const size_t depth = (size_t)6 * g->surf.pitch[l], chunk = depth < 32 ? depth : 32;
        void* zeros = nullptr;
        e = cudaMalloc(&zeros, n * n * 4 * chunk);
        if (e != cudaSuccess) break;
        cudaMemsetAsync(zeros, 0, n * n * 4 * chunk, dev->stream);
        for (size_t z0 = 0; z0 < depth && e == cudaSuccess; z0 += chunk) {
          cudaMemcpy3DParms z;
          memset(&z, 0, sizeof z);
          z.srcPtr = make_cudaPitchedPtr(zeros, n * 4, n, n);
          z.dstArray = arr;
          z.dstPos = make_cudaPos(0, 0, z0);
          z.extent = make_cudaExtent(n, n, depth - z0 < chunk ? depth - z0 : chunk);
          z.kind = cudaMemcpyDeviceToDevice;
          e = cudaMemcpy3DAsync(&z, dev->stream);
        }
        cudaStreamSynchronize(dev->stream);
        cudaFree(zeros);
      }
    }
    if (e == cudaSuccess) {
      cudaResourceDesc rd;
      memset(&rd, 0, sizeof rd);
      rd.resType = cudaResourceTypeMipmappedArray;
      rd.res.mipmap.mipmap = g->marr;
      cudaTextureDesc td;
      memset(&td, 0, sizeof td);
      td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeBorder;  // CLAMP_TO_BORDER, border (0,0,0,0): texture_3d.cpp:10-12
      td.filterMode = cudaFilterModeLinear;                                               // GL_LINEAR_MIPMAP_LINEAR: texture_3d.cpp:14
      td.mipmapFilterMode = cudaFilterModeLinear;
      td.readMode = cudaReadModeNormalizedFloat;
      td.normalizedCoords = 1;
      td.minMipmapLevelClamp = 0.0f;
      td.maxMipmapLevelClamp = (float)(levels - 2);
      e = cudaCreateTextureObject(&g->tex_lin, &rd, &td, nullptr);
      td.mipmapFilterMode = cudaFilterModePoint;   // same texels, one (the nearest) level per fetch
      if (e == cudaSuccess) e = cudaCreateTextureObject(&g->tex_one, &rd, &td, nullptr);
      td.filterMode = cudaFilterModePoint;         // the raw texel: what the software sampler (fp32 weights, rule R7) fetches
      td.readMode = cudaReadModeElementType;
      if (e == cudaSuccess) e = cudaCreateTextureObject(&g->tex_pt, &rd, &td, nullptr);
      g->tex_zs = (float)n_coarse / (float)(6 * pitch_coarse);
    }
  }
  if (e != cudaSuccess) {
    set_error("grid allocation failed: %s", cudaGetErrorString(e));
    vct_grid_destroy(g);
    return VCT_ERR_OOM;
  }
  *out = g;
  return vct_grid_clear(g);
}

int vct_grid_destroy(vct_grid_t* g) {
  if (!g) return VCT_OK;
  if (g->dev->peer_grid == g) vct_peer_disconnect(g->dev);
  cudaStreamSynchronize(g->dev->stream);
  cudaFree(g->base_buf[0]); cudaFree(g->base_buf[1]); cudaFree(g->tile_zero); cudaFree(g->tile_touched);
  for (int l = 0; l < VCT_MAX_LEVELS; l++) for (int d = 0; d < 6; d++) cudaFree(g->f16_lvl[l][d]);
  cudaFree(g->f16_table);
  cudaFree(g->rec3); cudaFree(g->rec_top); cudaFree(g->occb); cudaFree(g->mip_counters); cudaFree(g->sb_epoch);
  if (g->dev->vox_owner == g) g->dev->vox_owner = nullptr;
  for (int l = 0; l < 4; l++) cudaFree(g->occ[l]);
  cudaFree(g->occ_hi);
  cudaFree(g->docc_all);
  if (g->tex_lin) cudaDestroyTextureObject(g->tex_lin);
  if (g->tex_one) cudaDestroyTextureObject(g->tex_one);
  if (g->tex_pt) cudaDestroyTextureObject(g->tex_pt);
  for (int l = 0; l < VCT_MAX_LEVELS; l++) if (g->surf.s[l]) cudaDestroySurfaceObject(g->surf.s[l]);
  if (g->marr) cudaFreeMipmappedArray(g->marr);
  delete g;
  return VCT_OK;
}

int vct_grid_clear(vct_grid_t* g) {
  VCT_REQUIRE(g, "grid is null");
  vct_device* dev = g->dev;
  if (g->base_zero && !g->external) { g->dirty_z0 = g->dirty_z1 = 0; return VCT_OK; }   // nothing has been written since the last clear
  if (g->fmt == VCT_GRID_RGBA16F) {   // storage variant: always the dense clear
    VCT_CUDA(cudaMemsetAsync(g->base, 0, (size_t)g->R * g->R * g->R * 8, dev->stream));
  } else if (g->sparse_clear_ok && !g->external && dev->vox_owner == g) {
    // the non-zero words are exactly the occupied list of the last voxelization: zero those (and their tile flags)
    int rc = launch_sparse_clear(dev, g);
    if (rc) return rc;
  } else {
    VCT_CUDA(cudaMemsetAsync(g->base, 0, (size_t)g->R * g->R * g->R * 4, dev->stream));
    if (g->tile_touched) VCT_CUDA(cudaMemsetAsync(g->tile_touched, 0, (size_t)g->R * g->R * g->R / 2048, dev->stream));
  }
  g->sparse_clear_ok = false;
  g->flags_valid = g->tile_touched != nullptr && !g->external;
  g->base_zero = !g->external;
  g->dirty_z0 = g->dirty_z1 = 0;
  return VCT_OK;
}

int vct_grid_upload_base(vct_grid_t* g, const uint32_t* host) {
  VCT_REQUIRE(g && host, "null argument");
  VCT_REQUIRE(g->fmt == VCT_GRID_RGBA8, "RGBA8 grids only");
  g->untrack();
  VCT_CUDA(cudaMemcpyAsync(g->base, host, (size_t)g->R * g->R * g->R * 4, cudaMemcpyHostToDevice, g->dev->stream));
  VCT_CUDA(cudaStreamSynchronize(g->dev->stream));
  return VCT_OK;
}

int vct_grid_download_f16(vct_grid_t* g, int level, int dir, uint64_t* host) {
  VCT_REQUIRE(g && host, "null argument");
  VCT_REQUIRE(g->fmt == VCT_GRID_RGBA16F, "not an RGBA16F grid");
  VCT_REQUIRE(level >= 0 && level < g->levels && dir >= 0 && dir < 6, "bad level / direction");
  const size_t N = (size_t)(g->R >> level);
  const void* src = level == 0 ? (const void*)g->base : (const void*)g->f16_lvl[level][dir];
  VCT_CUDA(cudaMemcpyAsync(host, src, N * N * N * 8, cudaMemcpyDeviceToHost, g->dev->stream));
  VCT_CUDA(cudaStreamSynchronize(g->dev->stream));
  return VCT_OK;
}

int vct_grid_download(vct_grid_t* g, int level, int dir, uint32_t* host) {
  VCT_REQUIRE(g && host, "null argument");
  VCT_REQUIRE(g->fmt == VCT_GRID_RGBA8, "RGBA8 grids only (vct_grid_download_f16 for the fp16 variant)");
  VCT_REQUIRE(level >= 0 && level < g->levels, "bad level");
  VCT_REQUIRE(dir >= 0 && dir < 6, "bad direction");
  cudaStream_t s = g->dev->stream;
  const size_t N = (size_t)(g->R >> level);
  if (level == 0) {
    VCT_CUDA(cudaMemcpyAsync(host, g->base, N * N * N * 4, cudaMemcpyDeviceToHost, s));
  } else {
    // direction `dir` of the stacked mipmapped array (the only copy of levels >= 1)
    cudaArray_t arr;
    VCT_CUDA(cudaGetMipmappedArrayLevel(&arr, g->marr, (unsigned)(level - 1)));
    cudaMemcpy3DParms p;
    memset(&p, 0, sizeof p);
    p.srcArray = arr;
    p.srcPos = make_cudaPos(0, 0, (size_t)dir * g->surf.pitch[level]);
    p.dstPtr = make_cudaPitchedPtr(host, N * 4, N, N);
    p.extent = make_cudaExtent(N, N, N);
    p.kind = cudaMemcpyDeviceToHost;
    VCT_CUDA(cudaMemcpy3DAsync(&p, s));
  }
  VCT_CUDA(cudaStreamSynchronize(s));
  return VCT_OK;
}

int vct_grid_download_array(vct_grid_t* g, int level, int dir, uint32_t* host) {
  VCT_REQUIRE(g && level >= 1, "bad level (the array holds levels >= 1)");
  return vct_grid_download(g, level, dir, host);
}

size_t vct_grid_occupancy_words(const vct_grid_t* g, int level, int dilated) {
  if (!g || level < 0 || level >= g->levels) return 0;
  return dilated ? docc_words(g->R >> level) : occ_words(g->R >> level);
}

int vct_grid_download_occupancy(vct_grid_t* g, int level, int dilated, uint32_t* host) {
  VCT_REQUIRE(g && host, "null argument");
  VCT_REQUIRE(level >= 0 && level < g->levels, "bad level");
  VCT_CUDA(cudaMemcpyAsync(host, dilated ? g->docc[level] : g->occ[level], vct_grid_occupancy_words(g, level, dilated) * 4, cudaMemcpyDeviceToHost,
                           g->dev->stream));
  VCT_CUDA(cudaStreamSynchronize(g->dev->stream));
  return VCT_OK;
}

void* vct_grid_base_device_ptr(vct_grid_t* g) {
  if (!g || g->fmt != VCT_GRID_RGBA8) return nullptr;
  g->external = true;   // the caller may write level 0 (NCCL all-gather of the z-slabs): no sparse bookkeeping from here on
  g->untrack();
  return (void*)g->base;
}
size_t vct_grid_bytes(const vct_grid_t* g) { return g ? g->bytes : 0; }

// ------------------------------------------------------------------ target
int vct_target_create(vct_device_t* dev, int W, int H, vct_target_t** out) {
  VCT_REQUIRE(dev && out, "null argument");
  VCT_REQUIRE(W > 0 && H > 0 && W <= 16384 && H <= 16384, "bad frame size");
  vct_target_t_* t = new (std::nothrow) vct_target_t_();
  if (!t) { set_error("out of host memory"); return VCT_ERR_OOM; }
  t->dev = dev; t->W = W; t->H = H;
  size_t n = (size_t)W * H;
  cudaError_t e = cudaMalloc(&t->vis, n * 8);
  if (e == cudaSuccess) e = cudaMalloc(&t->world_pos, n * 12);
  if (e == cudaSuccess) e = cudaMalloc(&t->normal, n * 12);
  if (e == cudaSuccess) e = cudaMalloc(&t->material, n * 4);
  if (e == cudaSuccess) e = cudaMalloc(&t->frame, n * 4);
  if (e != cudaSuccess) {
    set_error("target allocation failed: %s", cudaGetErrorString(e));
    vct_target_destroy(t);
    return VCT_ERR_OOM;
  }
  int rc = launch_fill_u32(dev->stream, t->material, n, VCT_NO_TRIANGLE);
  if (rc) return rc;
  *out = t;
  return VCT_OK;
}

int vct_target_destroy(vct_target_t* t) {
  if (!t) return VCT_OK;
  if (t->dev->peer_target == t) vct_peer_disconnect(t->dev);
  cudaStreamSynchronize(t->dev->stream);
  cudaFree(t->vis); cudaFree(t->world_pos); cudaFree(t->normal); cudaFree(t->material); cudaFree(t->frame);
  cudaFree(t->cone_out); cudaFree(t->tile_list);
  t->pending_host = nullptr;
  if (t->copy_stream) { cudaStreamSynchronize(t->copy_stream); cudaStreamDestroy(t->copy_stream); }
  if (t->copy_gate) cudaEventDestroy(t->copy_gate);
  for (int i = 0; i < 2; i++) {
    cudaFree(t->snap[i]);
    if (t->snap_ready[i]) cudaEventDestroy(t->snap_ready[i]);
    if (t->copy_done[i]) cudaEventDestroy(t->copy_done[i]);
  }
  delete t;
  return VCT_OK;
}

int vct_target_download_frame(vct_target_t* t, uint32_t* host) {
  VCT_REQUIRE(t && host, "null argument");
  VCT_CUDA(cudaMemcpyAsync(host, t->frame, (size_t)t->W * t->H * 4, cudaMemcpyDeviceToHost, t->dev->stream));
  VCT_CUDA(cudaStreamSynchronize(t->dev->stream));
  return check_status(t->dev);   // the pixels were copied, but they come from a frame that dropped fragments / missed a peer
}

// Asynchronous read-back.  The finished frame is snapshotted on the device stream by a kernel (8 MB at 1080p, a few us), the
// snapshot is copied to the host on a second stream, and the call returns at once.  Two snapshots alternate; the device stream
// waits for the copy that last read the one it is about to overwrite.
// WHEN the copy runs matters: a device-to-host transfer in flight slows down the start of the next frame -- a chain of a dozen
// small kernels whose command fetches are PCIe reads that queue behind the posted writes of the transfer (measured: +83 us per
// frame for 8 MB, +5 us for 4 KB).  So the copy is not enqueued here: the next vct_render_frame on this target starts it
// together with its cone kernel (one 0.85 ms launch that needs nothing from the host), and vct_target_download_wait /
// a blocking download start it at once if no frame came in between.
static int start_pending_readback(vct_target_t_* t, cudaEvent_t gate) {
  if (!t->pending_host) return VCT_OK;
  const int b = t->pending_slot;
  const size_t bytes = (size_t)t->W * t->H * 4;
  VCT_CUDA(cudaStreamWaitEvent(t->copy_stream, t->snap_ready[b], 0));
  if (gate) VCT_CUDA(cudaStreamWaitEvent(t->copy_stream, gate, 0));
  VCT_CUDA(cudaMemcpyAsync(t->pending_host, t->snap[b], bytes, cudaMemcpyDeviceToHost, t->copy_stream));
  VCT_CUDA(cudaEventRecord(t->copy_done[b], t->copy_stream));
  t->pending_host = nullptr;
  return VCT_OK;
}

int vct_target_download_frame_async(vct_target_t* t, uint32_t* host, uint64_t* ticket) {
  VCT_REQUIRE(t && host && ticket, "null argument");
  cudaStream_t s = t->dev->stream;
  const size_t bytes = (size_t)t->W * t->H * 4;
  if (!t->copy_stream) {
    VCT_CUDA(cudaStreamCreateWithFlags(&t->copy_stream, cudaStreamNonBlocking));
    VCT_CUDA(cudaEventCreateWithFlags(&t->copy_gate, cudaEventDisableTiming));
    for (int i = 0; i < 2; i++) {
      VCT_CUDA(cudaMalloc(&t->snap[i], bytes));
      VCT_CUDA(cudaEventCreateWithFlags(&t->snap_ready[i], cudaEventDisableTiming));
      VCT_CUDA(cudaEventCreateWithFlags(&t->copy_done[i], cudaEventDisableTiming));
    }
  }
  int rc = start_pending_readback(t, nullptr);   // two read-backs without a frame in between: the older one goes now
  if (rc) return rc;
  const int b = (int)(t->n_async & 1);
  if (t->n_async >= 2) VCT_CUDA(cudaStreamWaitEvent(s, t->copy_done[b], 0));
  if ((rc = launch_copy_u32(s, t->snap[b], t->frame, (size_t)t->W * t->H))) return rc;
  VCT_CUDA(cudaEventRecord(t->snap_ready[b], s));
  t->pending_host = host;
  t->pending_slot = b;
  *ticket = ++t->n_async;
  return VCT_OK;
}

int vct_target_download_wait(vct_target_t* t, uint64_t ticket) {
  VCT_REQUIRE(t, "target is null");
  VCT_REQUIRE(ticket >= 1 && ticket <= t->n_async, "unknown ticket");
  if (t->pending_host && ticket == t->n_async) {   // its copy has not been started yet
    int rc = start_pending_readback(t, nullptr);
    if (rc) return rc;
  }
  // copy_done[b] always carries the newest copy out of snapshot b, issued at or after `ticket` on the same (ordered) copy stream
  VCT_CUDA(cudaEventSynchronize(t->copy_done[(ticket - 1) & 1]));
  return VCT_OK;
}

int vct_target_download_gbuffer(vct_target_t* t, uint32_t* tri_id, float* depth, float* world_pos, float* normal, uint32_t* material) {
  VCT_REQUIRE(t, "target is null");
  cudaStream_t s = t->dev->stream;
  size_t n = (size_t)t->W * t->H;
  if (tri_id) VCT_CUDA(cudaMemcpy2DAsync(tri_id, 4, t->vis, 8, 4, n, cudaMemcpyDeviceToHost, s));                      // low word
  if (depth) VCT_CUDA(cudaMemcpy2DAsync(depth, 4, (const uint32_t*)t->vis + 1, 8, 4, n, cudaMemcpyDeviceToHost, s));  // high word
  if (world_pos) VCT_CUDA(cudaMemcpyAsync(world_pos, t->world_pos, n * 12, cudaMemcpyDeviceToHost, s));
  if (normal) VCT_CUDA(cudaMemcpyAsync(normal, t->normal, n * 12, cudaMemcpyDeviceToHost, s));
  if (material) VCT_CUDA(cudaMemcpyAsync(material, t->material, n * 4, cudaMemcpyDeviceToHost, s));
  VCT_CUDA(cudaStreamSynchronize(s));
  return VCT_OK;
}

void* vct_target_frame_device_ptr(vct_target_t* t) { return t ? (void*)t->frame : nullptr; }

// ------------------------------------------------------------------ hot path
int vct_voxelize_reserve(vct_device_t* dev, uint64_t max_fragments) {
  VCT_REQUIRE(dev, "device is null");
  VCT_REQUIRE(max_fragments > 0 && max_fragments < 0xFFFFFFF0ull, "fragment capacity out of range");
  if (max_fragments <= dev->frag_capacity) return VCT_OK;
  VCT_CUDA(cudaStreamSynchronize(dev->stream));
  cudaFree(dev->frags); cudaFree(dev->fresh);
  dev->frags = nullptr; dev->fresh = nullptr; dev->frag_capacity = 0;
  dev->vox_owner = nullptr;   // the occupied list is gone: the next clear is dense
  VCT_CUDA(cudaMalloc(&dev->frags, max_fragments * sizeof(FragRec)));
  VCT_CUDA(cudaMalloc(&dev->fresh, max_fragments));
  dev->frag_capacity = max_fragments;
  return VCT_OK;
}

int vct_voxelize(vct_device_t* dev, vct_scene_t* sc, vct_grid_t* g, int z0, int z1) {
  VCT_REQUIRE(dev && sc && g, "null argument");
  VCT_REQUIRE(z0 >= 0 && z1 <= g->R && z0 <= z1, "bad z slab");
  // The level-0 word doubles as the head of the voxel's fragment list while the slab is being voxelized: content left in the slab
  // (an earlier vct_voxelize of an overlapping slab, vct_grid_upload_base) would be taken for list links.  The reference would
  // keep averaging into it (voxelize.frag:95-120); here the precondition of the header ("cleared") is enforced.
  VCT_REQUIRE(g->external || z0 >= g->dirty_z1 || z1 <= g->dirty_z0 || g->dirty_z0 >= g->dirty_z1,
              "the z slab overlaps voxels written since the last vct_grid_clear (voxelize each slab once per clear)");
  if (int rc = check_status(dev)) return rc;
  if (int rc = scene_ready(sc)) return rc;
  return launch_voxelize(dev, sc, g, z0, z1);
}

int vct_voxelize_stats(vct_device_t* dev, vct_voxel_stats_t* out) {
  VCT_REQUIRE(dev && out, "null argument");
  VCT_CUDA(cudaMemcpyAsync(dev->counters_host, dev->counters, CNT_TOTAL * sizeof(uint32_t), cudaMemcpyDeviceToHost, dev->stream));
  VCT_CUDA(cudaStreamSynchronize(dev->stream));
  out->items = dev->counters_host[CNT_ITEMS];
  out->fragments = dev->counters_host[CNT_FRAGS];
  out->occupied = dev->counters_host[CNT_OCCUPIED];
  out->max_per_voxel = dev->counters_host[CNT_MAXLIST];
  out->capacity = dev->frag_capacity;
  return check_status(dev);   // VCT_ERR_OVERFLOW (arena grown) if this voxelization dropped fragments
}

int vct_voxelize_set_accum_mode(vct_device_t* dev, int mode) {
  VCT_REQUIRE(dev, "device is null");
  VCT_REQUIRE(mode == VCT_ACCUM_ORDERED || mode == VCT_ACCUM_FIXED_POINT, "accumulation mode must be VCT_ACCUM_ORDERED or VCT_ACCUM_FIXED_POINT");
  dev->accum_mode = mode;
  return VCT_OK;
}

int vct_debug_set(vct_device_t* dev, int key, int value) {
  VCT_REQUIRE(dev, "device is null");
  switch (key) {
    case VCT_DEBUG_MIP_DENSE: dev->debug_mip_dense = value != 0; return VCT_OK;
    case VCT_DEBUG_CONE_VARIANT: VCT_REQUIRE(value >= -1 && value <= 3, "cone variant must be -1..3"); dev->debug_cone_variant = value; return VCT_OK;
    case VCT_DEBUG_CONE_GRID: dev->debug_cone_grid = value != 0; return VCT_OK;
    case VCT_DEBUG_CONE_CTAS_PER_SM: VCT_REQUIRE(value >= 0 && value <= 16, "CTAs per SM out of range"); dev->cone_ctas_per_sm = value; return VCT_OK;
    case VCT_DEBUG_SMALL_LIMIT: VCT_REQUIRE(value >= -1 && value <= 1024, "small-triangle limit out of range"); dev->debug_small_limit = value; return VCT_OK;
    case VCT_DEBUG_PEER_REPLICATE:
      VCT_REQUIRE(value >= -1 && value <= 1, "peer replicate must be -1 (automatic), 0 or 1");
      VCT_REQUIRE(dev->peers.nranks <= 1, "set before vct_peer_connect");
      dev->peer_replicate_force = value;
      return VCT_OK;
    case VCT_DEBUG_TRACE_LOW_PRIORITY: dev->trace_low_priority = value != 0; return VCT_OK;
    case VCT_DEBUG_CONE_RESERVE_SMS:
      VCT_REQUIRE(value >= 0 && value < dev->prop.multiProcessorCount, "reserved SM count out of range");
      dev->cone_reserved_sms = value;
      return VCT_OK;
  }
  set_error("vct_debug_set: unknown key %d", key);
  return VCT_ERR_INVALID;
}

int vct_mipmap(vct_device_t* dev, vct_grid_t* g) {
  VCT_REQUIRE(dev && g, "null argument");
  return launch_mipmap(dev, g);
}

int vct_gbuffer(vct_device_t* dev, vct_scene_t* sc, const float view[16], const float proj[16], vct_target_t* t) {
  VCT_REQUIRE(dev && sc && view && proj && t, "null argument");
  if (int rc = scene_ready(sc)) return rc;
  return launch_gbuffer(dev, sc, view, proj, t);
}

int vct_cone_trace(vct_device_t* dev, vct_scene_t* sc, vct_grid_t* g, const float view[16], const vct_trace_params_t* p, vct_target_t* t) {
  VCT_REQUIRE(dev && sc && g && view && p && t, "null argument");
  if (int rc = scene_ready(sc)) return rc;
  return launch_cone_trace(dev, sc, g, view, p, t, false);
}

int vct_cone_trace_count(vct_device_t* dev, vct_scene_t* sc, vct_grid_t* g, const float view[16], const vct_trace_params_t* p, vct_target_t* t,
                         vct_trace_stats_t* out) {
  VCT_REQUIRE(dev && sc && g && view && p && t && out, "null argument");
  int rc = scene_ready(sc);
  if (rc) return rc;
  rc = launch_cone_trace(dev, sc, g, view, p, t, true);
  if (rc) return rc;
  unsigned long long h[8];
  VCT_CUDA(cudaMemcpyAsync(h, dev->counters + 16, sizeof h, cudaMemcpyDeviceToHost, dev->stream));
  VCT_CUDA(cudaStreamSynchronize(dev->stream));
  out->samples_diffuse = h[0]; out->samples_shadow = h[1]; out->samples_specular = h[2]; out->samples_refraction = h[3];
  out->shaded_pixels = h[4];
  return VCT_OK;
}

// One rank's share of a frame after vct_peer_connect.  Buffer-reuse argument (e = frame number, b = e & 1):
//  * this rank pushes frame e into base_buf[b] of every peer.  A peer cleared that buffer during ITS frame e-1 (below:
//    "clear the next buffer") before it published PUSHED(e-1), and this rank waited for PUSHED(e-1) of every peer in
//    its frame e-1 -> the clear happened before the push.
//  * the buffer cleared here, base_buf[b^1], was last written by pushes of frame e-1, all complete once every PUSHED(e-1)
//    flag was seen (frame e-1 of this rank), and last read by this rank's own trace of frame e-1 (stream order).
//  * a peer's frame buffer receives tiles of frame e only after that peer published PUSHED(e), i.e. after everything it
//    had enqueued for frame e-1 (including a frame download) in stream order.
// the trace of one rank's screen tiles (+ their push to the root) and, on the root, the wait for everybody's tiles
static int sharded_trace(vct_device_t* dev, vct_scene_t* sc, vct_grid_t* g, vct_target_t* t, const float view[16], const vct_trace_params_t* prm,
                         const PeerView& pv, uint32_t epoch) {
  cudaStream_t s = dev->stream;
  int rc;
  if (dev->trace_low_priority) {   // frames in flight: see vct_render_frame
    VCT_CUDA(cudaEventRecord(dev->ev_front, s));
    VCT_CUDA(cudaStreamWaitEvent(dev->stream3, dev->ev_front, 0));
    dev->stream = dev->stream3;
  }
  rc = launch_cone_trace(dev, sc, g, view, prm, t, false, &pv, 2);                           // + push of the tiles to the root
  if (dev->trace_low_priority) {
    dev->stream = s;
    if (rc) return rc;
    VCT_CUDA(cudaEventRecord(dev->ev_trace, dev->stream3));
    VCT_CUDA(cudaStreamWaitEvent(s, dev->ev_trace, 0));
  }
  if (rc) return rc;
  // the root's wait for the other ranks' tiles sits on this pipeline's own stream: it holds back what follows on this device object
  // (a download, its next frame), not the shared trace stream
  if (pv.frame_root < 0 || pv.frame_root == pv.rank)
    if ((rc = launch_peer_wait(dev, PEER_FLAG_FRAME, epoch))) return rc;                    // every tile has arrived
  VCT_CUDA(cudaEventRecord(dev->ev[5], s));
  dev->have_timings = true;
  dev->gbuffer_overlapped = true;
  return VCT_OK;
}

static int render_frame_sharded(vct_device_t* dev, vct_scene_t* sc, vct_grid_t* g, vct_target_t* t, const float view[16], const float proj[16],
                                const vct_trace_params_t* p) {
  cudaStream_t s = dev->stream;
  const uint32_t epoch = ++dev->peer_epoch;
  const int b = (int)(epoch & 1u);
  PeerView pv = dev->peers;
  pv.epoch = epoch;
  for (int r = 0; r < pv.nranks; r++) { pv.base[r] = dev->peer_base_all[b][r]; pv.touched[r] = dev->peer_touched_all[b][r]; }
  // A small scene (the reference's Cornell box: 1 k triangles) voxelizes in ~70 us of launch and dependency latency whatever the slab:
  // sharding it saves nothing and adds the peers' flag waits to every rank's front half.  Every rank then voxelizes the whole scene into
  // its own grid (no voxel exchange, the single-GPU sparse clear / sparse mip bookkeeping), and only the screen tiles are split.
  // Decided once per connection from the scene of its first frame (the same on every rank).
  if (dev->peer_replicate < 0) {
    dev->peer_replicate = dev->peer_replicate_force >= 0 ? dev->peer_replicate_force : (sc->n_tris <= 16384u ? 1 : 0);
    if (dev->peer_replicate == 1) {   // nobody stores into this grid: back to the tracked single-GPU state (the first clear is dense)
      g->base = g->base_buf[0];
      g->external = false;
      g->peer_touched = nullptr;
      g->untrack();
    }
  }
  if (dev->peer_replicate == 1) {
    cudaStream_t s = dev->stream;
    vct_trace_params_t prm = *p;
    prm.tile_rank = pv.rank; prm.tile_nranks = pv.nranks;
    int rc;
    VCT_CUDA(cudaEventRecord(dev->ev[0], s));
    VCT_CUDA(cudaEventRecord(dev->ev_fork, s));
    if ((rc = vct_grid_clear(g))) return rc;
    VCT_CUDA(cudaEventRecord(dev->ev[1], s));
    if ((rc = launch_voxelize(dev, sc, g, 0, g->R))) return rc;
    VCT_CUDA(cudaEventRecord(dev->ev[2], s));
    VCT_CUDA(cudaStreamWaitEvent(dev->stream2, dev->ev_fork, 0));
    dev->stream = dev->stream2;
    VCT_CUDA(cudaEventRecord(dev->ev_g0, dev->stream2));
    const bool fused_list = prm.view_voxel_dir >= 7;
    // PUSHED(e) of rank r now only says "r has finished everything it had queued before frame e" (the fork event): its frame buffer may
    // receive tiles of e.  Sent from the second stream: a system fence + N remote stores that the critical path need not wait for.
    rc = launch_peer_signal(dev, pv, PEER_FLAG_PUSHED);
    if (!rc) rc = launch_gbuffer(dev, sc, view, proj, t, pv.rank, pv.nranks, fused_list);
    if (!rc && !fused_list) rc = launch_cone_trace(dev, sc, g, view, &prm, t, false, &pv, 1);
    dev->stream = s;
    if (rc) return rc;
    VCT_CUDA(cudaEventRecord(dev->ev_g1, dev->stream2));
    VCT_CUDA(cudaEventRecord(dev->ev_join, dev->stream2));
    if ((rc = launch_mipmap(dev, g))) return rc;
    VCT_CUDA(cudaEventRecord(dev->ev[3], s));
    VCT_CUDA(cudaStreamWaitEvent(s, dev->ev_join, 0));
    if ((rc = launch_peer_wait(dev, PEER_FLAG_PUSHED, epoch))) return rc;   // the destination frame(s) are free (long since: the flags went out at the head of the frame)
    VCT_CUDA(cudaEventRecord(dev->ev[4], s));
    return sharded_trace(dev, sc, g, t, view, &prm, pv, epoch);
  }
  g->base = g->base_buf[b];
  // every rank exported tile flags (same grid size everywhere: all or none); a mip tile is 8 slices deep and must belong to ONE slab
  // Decided once per connection, from the scene of its first frame (the same on every rank): the un-push stores nranks words + flags per
  // occupied voxel over NVLink, which beats a 4 R^3-byte local clear + a dense mip build only while the scene is sparse.  Measured at
  // N = 8: 256^3 / 1 k triangles 340 -> 326 us per frame, but 1024^3 / 4 M triangles (23 M occupied voxels) clear 0.60 -> 1.36 ms.
  if (dev->peer_sparse_mode < 0) dev->peer_sparse_mode = sc->n_tris <= 65536u ? 1 : 0;
  const bool sparse = dev->peer_sparse_mode == 1 && pv.touched[pv.rank] != nullptr && g->R % (8 * pv.nranks) == 0;
  if (!sparse) for (int r = 0; r < pv.nranks; r++) pv.touched[r] = nullptr;
  if (sparse) {
    if (dev->frag_capacity == 0)
      if (int rcr = vct_voxelize_reserve(dev, 1u << 20)) return rcr;
    if (int rcl = ensure_pushed_lists(dev)) return rcl;
    pv.pushed = dev->pushed_list[b]; pv.pushed_n = dev->pushed_n + (epoch & 3u); pv.pushed_capacity = (uint32_t)dev->pushed_capacity;
  }
  g->peer_touched = sparse ? pv.touched[pv.rank] : nullptr;
  const int z0 = (int)((long long)pv.rank * g->R / pv.nranks), z1 = (int)((long long)(pv.rank + 1) * g->R / pv.nranks);
  vct_trace_params_t prm = *p;
  prm.tile_rank = pv.rank; prm.tile_nranks = pv.nranks;
  int rc;
  VCT_CUDA(cudaEventRecord(dev->ev[0], s));
  VCT_CUDA(cudaEventRecord(dev->ev_fork, s));
  if (sparse) {
    int logR = 0;
    while ((1 << logR) < g->R) logR++;
    if ((rc = launch_peer_unpush(dev, pv, logR))) return rc;   // this frame's buffer: zero what this rank stored into it two frames ago, everywhere
  } else {
    VCT_CUDA(cudaMemsetAsync(g->base_buf[b ^ 1], 0, (size_t)g->R * g->R * g->R * 4, s));   // clear the NEXT frame's buffer
  }
  VCT_CUDA(cudaEventRecord(dev->ev[1], s));
  if ((rc = launch_voxelize(dev, sc, g, z0, z1, &pv))) return rc;   // + push of the slab to the peers
  VCT_CUDA(cudaEventRecord(dev->ev[2], s));
  // visibility of this rank's tiles + their live-tile list on the second stream, beside the voxelization, the wait for the peers' slabs
  // and the mip build (as in the single-GPU frame)
  VCT_CUDA(cudaStreamWaitEvent(dev->stream2, dev->ev_fork, 0));
  dev->stream = dev->stream2;
  VCT_CUDA(cudaEventRecord(dev->ev_g0, dev->stream2));
  const bool fused_list = prm.view_voxel_dir >= 7;   // the G-buffer resolve builds the live-tile list (the debug view shades every tile: no list)
  rc = launch_gbuffer(dev, sc, view, proj, t, pv.rank, pv.nranks, fused_list);
  if (!rc && !fused_list) rc = launch_cone_trace(dev, sc, g, view, &prm, t, false, &pv, 1);
  dev->stream = s;
  if (rc) return rc;
  VCT_CUDA(cudaEventRecord(dev->ev_g1, dev->stream2));
  VCT_CUDA(cudaEventRecord(dev->ev_join, dev->stream2));
  if ((rc = launch_peer_wait(dev, PEER_FLAG_PUSHED, epoch))) return rc;                     // every slab has arrived
  if ((rc = launch_mipmap(dev, g))) return rc;
  VCT_CUDA(cudaEventRecord(dev->ev[3], s));
  VCT_CUDA(cudaStreamWaitEvent(s, dev->ev_join, 0));
  VCT_CUDA(cudaEventRecord(dev->ev[4], s));
  return sharded_trace(dev, sc, g, t, view, &prm, pv, epoch);
}

int vct_render_frame(vct_device_t* dev, vct_scene_t* sc, vct_grid_t* g, vct_target_t* t, const float view[16], const float proj[16],
                     const vct_trace_params_t* p) {
  VCT_REQUIRE(dev && sc && g && t && view && proj && p, "null argument");
  if (int rc0 = check_status(dev)) return rc0;   // an earlier frame overflowed the fragment arena (grown now: re-issue) or missed a peer
  if (int rc0 = scene_ready(sc)) return rc0;   // the frame stream waits for the scene uploads in flight (they run on their own stream)
  if (dev->peers.nranks > 1 && dev->peer_grid == g && dev->peer_target == t) return render_frame_sharded(dev, sc, g, t, view, proj, p);
  cudaStream_t s = dev->stream;
  int rc;
  VCT_CUDA(cudaEventRecord(dev->ev[0], s));
  // The G-buffer pass needs the scene and the camera, not the voxel grid: it runs on a second stream beside
  // clear + voxelize + mip (chains of small, latency-bound kernels that leave most SMs idle) and joins before the trace.
  VCT_CUDA(cudaEventRecord(dev->ev_fork, s));               // everything queued so far (scene uploads, the previous frame's read-back)
  if ((rc = vct_grid_clear(g))) return rc;                  // the critical path is submitted first, the G-buffer launches fill in behind it
  VCT_CUDA(cudaEventRecord(dev->ev[1], s));
  if ((rc = launch_voxelize(dev, sc, g, 0, g->R))) return rc;
  VCT_CUDA(cudaEventRecord(dev->ev[2], s));
  VCT_CUDA(cudaStreamWaitEvent(dev->stream2, dev->ev_fork, 0));
  dev->stream = dev->stream2;
  VCT_CUDA(cudaEventRecord(dev->ev_g0, dev->stream2));
  const bool fused_list = p->view_voxel_dir >= 7;   // the G-buffer resolve builds the live-tile list (the debug view shades every tile: no list)
  rc = launch_gbuffer(dev, sc, view, proj, t, 0, 1, fused_list);
  if (!rc && !fused_list) rc = launch_cone_trace(dev, sc, g, view, p, t, false, nullptr, 1);
  dev->stream = s;
  if (rc) return rc;
  VCT_CUDA(cudaEventRecord(dev->ev_g1, dev->stream2));
  VCT_CUDA(cudaEventRecord(dev->ev_join, dev->stream2));
  if ((rc = launch_mipmap(dev, g))) return rc;
  VCT_CUDA(cudaEventRecord(dev->ev[3], s));
  VCT_CUDA(cudaStreamWaitEvent(s, dev->ev_join, 0));
  VCT_CUDA(cudaEventRecord(dev->ev[4], s));                 // ev[3]..ev[4] = what is left of the G-buffer pass after the mip build
  if (t->pending_host) {   // the previous frame's read-back crosses PCIe while the cone kernel runs (see vct_target_download_frame_async)
    VCT_CUDA(cudaEventRecord(t->copy_gate, s));
    if ((rc = start_pending_readback(t, t->copy_gate))) return rc;
  }
  if (dev->trace_low_priority) {   // frames in flight: cones + shade on the low-priority stream, behind another pipeline's front half
    VCT_CUDA(cudaEventRecord(dev->ev_front, s));
    VCT_CUDA(cudaStreamWaitEvent(dev->stream3, dev->ev_front, 0));
    dev->stream = dev->stream3;
  }
  rc = launch_cone_trace(dev, sc, g, view, p, t, false, nullptr, 2);
  if (dev->trace_low_priority) {
    dev->stream = s;
    if (rc) return rc;
    VCT_CUDA(cudaEventRecord(dev->ev_trace, dev->stream3));
    VCT_CUDA(cudaStreamWaitEvent(s, dev->ev_trace, 0));
  }
  if (rc) return rc;
  VCT_CUDA(cudaEventRecord(dev->ev[5], s));
  dev->have_timings = true;
  dev->gbuffer_overlapped = true;
  return VCT_OK;
}

int vct_debug_frame_events(vct_device_t* dev, vct_device_t* ref, float out_ms[8]) {
  VCT_REQUIRE(dev && ref && out_ms, "null argument");
  VCT_REQUIRE(dev->have_timings && ref->have_timings, "no frame has been rendered");
  VCT_CUDA(cudaEventSynchronize(dev->ev[5]));
  VCT_CUDA(cudaEventSynchronize(ref->ev[5]));
  for (int i = 0; i < 8; i++) {
    out_ms[i] = 0.0f;
    if (cudaEventElapsedTime(&out_ms[i], ref->ev[0], dev->ev[i]) != cudaSuccess) { out_ms[i] = -1.0f; cudaGetLastError(); }
  }
  return VCT_OK;
}

int vct_last_frame_timings(vct_device_t* dev, float out_ms[8]) {
  VCT_REQUIRE(dev && out_ms, "null argument");
  VCT_REQUIRE(dev->have_timings, "no frame has been rendered");
  VCT_CUDA(cudaEventSynchronize(dev->ev[5]));
  for (int i = 0; i < 5; i++) VCT_CUDA(cudaEventElapsedTime(&out_ms[i], dev->ev[i], dev->ev[i + 1]));
  VCT_CUDA(cudaEventElapsedTime(&out_ms[5], dev->ev[0], dev->ev[5]));
  out_ms[6] = out_ms[7] = 0.0f;
  if (dev->gbuffer_overlapped) VCT_CUDA(cudaEventElapsedTime(&out_ms[7], dev->ev_g0, dev->ev_g1));
  if (cudaEventQuery(dev->ev[7]) == cudaSuccess && cudaEventElapsedTime(&out_ms[6], dev->ev[6], dev->ev[7]) != cudaSuccess) {
    out_ms[6] = 0.0f;
    cudaGetLastError();
  }
  return VCT_OK;
}

}  // extern "C"
