/*
 * vct_c.h -- the C ABI of libvct_cuda.so: the drop-in boundary of the B200-native voxel cone
 * tracing hot path (voxelize -> anisotropic mip build -> G-buffer -> cone trace).
 *
 * The reference (latencyhiding/voxel_cone_tracing) has no FFI/plugin system: its hot path sits
 * behind `struct Renderer` (src/renderer.h:122-196) and is executed by OpenGL.  Every entry point
 * below names the reference interface it replaces.  The C++ host layer (include/vct/renderer.h,
 * texture_3d.h, device.h) and the Python/ctypes harness both sit on top of this file.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 (VCT_OK) or a negative
 * error code and never throws; vct_last_error() returns a thread-local message; host pointers are
 * borrowed for the duration of the call; all work is enqueued on the device's stream and is
 * asynchronous unless the function copies results to the host.  There is NO CPU fallback: every
 * call fails with VCT_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef VCT_C_H
#define VCT_C_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VCT_OK 0
#define VCT_ERR_INVALID (-1) /* bad argument                                   */
#define VCT_ERR_CUDA (-2)    /* CUDA runtime / driver error, or no usable GPU   */
#define VCT_ERR_OOM (-3)     /* device or pinned-host allocation failed         */
#define VCT_ERR_OVERFLOW (-4) /* fragment arena too small; see vct_voxelize     */

#define VCT_MAX_POINT_LIGHTS 10 /* shader/voxelize.frag:40 */
#define VCT_NO_TRIANGLE 0xFFFFFFFFu

typedef struct vct_device vct_device_t; /* replaces the (empty) class Device, src/device.h:32-38 */
typedef struct vct_scene vct_scene_t;   /* replaces m_models/m_draw_objs/m_materials/m_draw_queue/m_point_lights, src/renderer.h:168-191 */
typedef struct vct_grid vct_grid_t;     /* replaces m_voxel_maps[6], src/renderer.h:180 (six GL 3-D textures) */
typedef struct vct_target_t_ vct_target_t; /* replaces the default framebuffer + depth buffer, src/renderer.cpp:355-361 */
typedef struct vct_tex3d vct_tex3d_t;   /* replaces one GLuint 3-D texture of src/texture_3d.h:6-10 */

/* vert_data_t, src/renderer.cpp:24-35 */
typedef struct { float pos[3]; float norm[3]; float uv[2]; } vct_vertex_t;
/* draw_obj_t + model_t::model_matrix, src/renderer.h:57-79; model is column-major (glm::mat4) */
typedef struct {
  uint32_t first_index, index_count, vertex_base, material;
  float model[16];
} vct_draw_t;
/* point_light_t, src/renderer.h:87-92 */
typedef struct { float position[3]; float color[3]; float intensity; } vct_point_light_t;
/* material_data_t, src/renderer.h:94-120 (128 bytes, std140-compatible) */
typedef struct {
  float ambient[4], diffuse[4], specular[4], transmittance[4];
  float emission[3];
  float shininess, ior, dissolve;
  int32_t illum;
  float roughness, metallic, sheen, clearcoat_thickness, clearcoat_roughness, anisotropy, anisotropy_rotation;
  float pad[2];
} vct_material_t;

/* uniforms of Renderer::visualize, src/renderer.cpp:365-374, + the multi-GPU tile split */
typedef struct {
  int32_t enable_direct, enable_diffuse, enable_specular, enable_shadow; /* set_rendering_phases, renderer.cpp:195-201 */
  int32_t view_voxel_dir; /* set_voxel_view_dir, renderer.cpp:203-207; >= 7 = shade normally */
  float view_voxel_lod;
  int32_t n_diffuse_cones; /* 9 = reference (voxel_cone_tracing.frag:153-165); 5 = normal + 4 side cones (BASELINE.json config 1);
                            * 16 = normal + 5 at 30 deg + 10 at 60 deg, aperture 2 tan 15 deg (config 5); non-reference variants */
  int32_t tile_rank, tile_nranks; /* this call shades the 32x32 screen tiles (tx, ty) with (tx + k ty) % tile_nranks == tile_rank, k = 3 (5 or 7 when 3 / 15 divides tile_nranks) */
  int32_t sampler; /* VCT_SAMPLER_*: how textureLod is evaluated */
} vct_trace_params_t;

#define VCT_SAMPLER_FP32 0 /* software trilinear + mip-linear with fp32 weights (oracle rule R7) */
#define VCT_SAMPLER_TEX 1  /* levels >= 1 filtered by the texture units (8-bit weights), level 0 in software */

typedef struct {
  uint64_t fragments;      /* fragments folded into the grid                    */
  uint64_t occupied;       /* voxels written                                    */
  uint64_t items;          /* 8x8 raster work items                             */
  uint64_t capacity;       /* fragment arena capacity                           */
  uint64_t max_per_voxel;  /* longest per-voxel fragment list                   */
} vct_voxel_stats_t;

typedef struct {
  uint64_t shaded_pixels;
  uint64_t samples_diffuse, samples_shadow, samples_specular, samples_refraction;
} vct_trace_stats_t;

/* ---- device (class Device, src/device.h:32-38; GL context creation src/main.cpp:55-76) ---- */
int vct_device_create(int cuda_ordinal, vct_device_t** out);
int vct_device_destroy(vct_device_t* dev);
int vct_device_sync(vct_device_t* dev);
void* vct_device_stream(vct_device_t* dev); /* cudaStream_t, for interop (torch / NCCL) */
const char* vct_last_error(void);
const char* vct_version(void);

/* ---- scene state (Renderer::load_model upload renderer.cpp:559-600, add_material/upload_material_data
 *      :643-665, queue_model :209-218, queue_point_light :220-223, set_grid_size :190-193) ---- */
int vct_scene_create(vct_device_t* dev, vct_scene_t** out);
int vct_scene_destroy(vct_scene_t* sc);
int vct_scene_set_geometry(vct_scene_t* sc, const vct_vertex_t* verts, uint32_t n_verts, const uint32_t* indices, uint32_t n_indices);
int vct_scene_set_materials(vct_scene_t* sc, const vct_material_t* mats, uint32_t n_mats);
int vct_scene_set_draws(vct_scene_t* sc, const vct_draw_t* draws, uint32_t n_draws);
int vct_scene_set_lights(vct_scene_t* sc, const vct_point_light_t* lights, uint32_t n_lights);
int vct_scene_set_cube_size(vct_scene_t* sc, float cube_size);

/* ---- voxel grid = the six directional textures (set_grid_resolution renderer.cpp:178-188,
 *      create_tex_3d texture_3d.cpp:3-25).  Level 0 is stored once (the reference writes the
 *      same value to all six, voxelize.frag:159-160); levels 1.. hold 6 directions per texel. ---- */
int vct_grid_create(vct_device_t* dev, int resolution, int levels, vct_grid_t** out);
/* Storage variant of BASELINE.json config 5 ("fp16 RGBA grid + full mip chain"; NOT the reference's format, which is RGBA8 with 7 levels:
 * texture_3d.cpp:3-25, renderer.cpp:186).  VCT_GRID_RGBA16F: every texel of every level is four IEEE halves (8 bytes); voxels hold the mean
 * fragment colour (fixed-point accumulation, see VCT_ACCUM_FIXED_POINT), the mip chain blends in fp32 and rounds to half, the cone tracer
 * filters in fp32 (software sampler; the texture-unit path and the multi-GPU exchange are RGBA8 only).  Any `levels` up to log2(R)+1. */
#define VCT_GRID_RGBA8 0
#define VCT_GRID_RGBA16F 1
int vct_grid_create_ex(vct_device_t* dev, int resolution, int levels, int format, vct_grid_t** out);
int vct_grid_download_f16(vct_grid_t* g, int level, int dir, uint64_t* host); /* (R >> level)^3 texels of four halves, R in the low 16 bits */
int vct_grid_destroy(vct_grid_t* g);
int vct_grid_clear(vct_grid_t* g);                                        /* clear_tex_3d x6, renderer.cpp:320-321 */
int vct_grid_upload_base(vct_grid_t* g, const uint32_t* host_rgba8);      /* R^3 texels, [z][y][x] */
int vct_grid_download(vct_grid_t* g, int level, int dir, uint32_t* host); /* glGetTexImage equivalent; level 0 ignores dir */
int vct_grid_download_array(vct_grid_t* g, int level, int dir, uint32_t* host); /* level >= 1; same as vct_grid_download (levels >= 1 live only in the mipmapped CUDA array the texture units sample) */
/* occupancy bit masks written by vct_mipmap and read by the cone tracer to skip all-zero filter footprints (no reference
 * counterpart; inspection only).  dilated = 0: bit (z*N + y)*N + x = texel non-zero in any direction; dilated = 1: volume of
 * (N+1)^3 bits in rows of (N+32)/32 words, bit (x+1,y+1,z+1) = any texel of [x,x+1]x[y,y+1]x[z,z+1] non-zero. */
size_t vct_grid_occupancy_words(const vct_grid_t* g, int level, int dilated);
int vct_grid_download_occupancy(vct_grid_t* g, int level, int dilated, uint32_t* host_words);
void* vct_grid_base_device_ptr(vct_grid_t* g);                            /* device pointer of level 0 (in-place NCCL allgather of z-slabs) */
size_t vct_grid_bytes(const vct_grid_t* g);

/* ---- render target: visibility + G-buffer + RGBA8 colour ---- */
int vct_target_create(vct_device_t* dev, int width, int height, vct_target_t** out);
int vct_target_destroy(vct_target_t* t);
int vct_target_download_frame(vct_target_t* t, uint32_t* host_rgba8);     /* row 0 = bottom (GL window coords) */
/* Asynchronous read-back (the SwapBuffers of this build, src/main.cpp:387): snapshots the finished frame in stream order and
 * copies it to `host_rgba8` (pinned memory for a truly asynchronous copy) on a second stream; returns at once with a ticket so
 * that the next vct_render_frame overlaps the transfer.  vct_target_download_wait blocks until that ticket's copy has landed. */
int vct_target_download_frame_async(vct_target_t* t, uint32_t* host_rgba8, uint64_t* ticket);
int vct_target_download_wait(vct_target_t* t, uint64_t ticket);
int vct_target_download_gbuffer(vct_target_t* t, uint32_t* tri_id, float* depth, float* world_pos3, float* normal3, uint32_t* material);
void* vct_target_frame_device_ptr(vct_target_t* t);

/* ---- the hot path ---- */
/* Renderer::voxelize() draw part (renderer.cpp:323-347 + voxelize.vert/geom/frag): folds every
 * fragment whose voxel z lies in [z0,z1) into level 0 with the reference's RGBA8 running average,
 * in the canonical order (draw, triangle, pixel row, pixel column).  The grid must have been
 * cleared.  VCT_ERR_OVERFLOW is reported by vct_voxelize_stats if the fragment arena was too small
 * (the arena is then grown; re-run the frame). */
int vct_voxelize(vct_device_t* dev, vct_scene_t* sc, vct_grid_t* g, int z0, int z1);
int vct_voxelize_reserve(vct_device_t* dev, uint64_t max_fragments);
int vct_voxelize_stats(vct_device_t* dev, vct_voxel_stats_t* out);        /* synchronises */
/* How the fragments of one voxel are combined (BASELINE.json north_star: "deterministic integer or fixed-point atomic accumulation path").
 *   VCT_ACCUM_ORDERED      (default) the reference's imageAtomicRGBA8Avg running average (voxelize.frag:95-120: 7-bit colour, 4-bit count that
 *                          wraps at 16) applied in canonical fragment order (draw, triangle, row, column): bit-exact against the oracle.
 *   VCT_ACCUM_FIXED_POINT  NON-REFERENCE variant: every fragment adds (uint)(colour * 255 + 0.5) per channel to 64-bit integer accumulators
 *                          (atomicAdd: order independent, so deterministic without sorting), the voxel stores the rounded mean in all 8 bits.
 *                          Differs from the ordered result by at most 3/255 per channel on the benchmark scenes (profiles/); no list walk,
 *                          no sort in the resolve pass. */
#define VCT_ACCUM_ORDERED 0
#define VCT_ACCUM_FIXED_POINT 1
int vct_voxelize_set_accum_mode(vct_device_t* dev, int mode);
/* Renderer::filter() (renderer.cpp:283-314 + mipmap.comp) */
int vct_mipmap(vct_device_t* dev, vct_grid_t* g);
/* vertex + raster + depth part of Renderer::visualize() (renderer.cpp:355-390, voxel_cone_tracing.vert) */
int vct_gbuffer(vct_device_t* dev, vct_scene_t* sc, const float view[16], const float proj[16], vct_target_t* t);
/* fragment part of Renderer::visualize() (voxel_cone_tracing.frag) */
int vct_cone_trace(vct_device_t* dev, vct_scene_t* sc, vct_grid_t* g, const float view[16], const vct_trace_params_t* p, vct_target_t* t);
/* same, instrumented: counts trace_cone loop iterations (untimed build of the kernel) */
int vct_cone_trace_count(vct_device_t* dev, vct_scene_t* sc, vct_grid_t* g, const float view[16], const vct_trace_params_t* p, vct_target_t* t,
                         vct_trace_stats_t* out);
/* Renderer::render() (renderer.cpp:392-405): clear + voxelize + mip + G-buffer + trace, one stream-ordered sequence */
int vct_render_frame(vct_device_t* dev, vct_scene_t* sc, vct_grid_t* g, vct_target_t* t, const float view[16], const float proj[16],
                     const vct_trace_params_t* p);

/* per-stage device timings (CUDA events on the device stream) of the last vct_render_frame, in ms:
 * [0] clear [1] voxelize [2] mipmap [3] gbuffer, the part on the critical path (single GPU: the pass runs on a second stream
 * beside [0]-[2], this is what is left of it after the mip build) [4] trace (tile list + cones + shade) [5] total
 * [6] cone kernel alone [7] the G-buffer pass itself on its own stream (0 when it ran in line); synchronises */
int vct_last_frame_timings(vct_device_t* dev, float out_ms[8]);
/* measurement: the event times of dev's last frame relative to the START of ref's last frame (two device objects of one process: frames
 * in flight), ms: [0] frame start [1] clear done [2] voxelize done [3] mip done [4] front half done (G-buffer joined) [5] frame done
 * [6] cone kernel start [7] cone kernel end; -1 where an event was not recorded.  Synchronises. */
int vct_debug_frame_events(vct_device_t* dev, vct_device_t* ref, float out_ms[8]);

/* measurement / test / tuning switches (replace the environment variables of round 1; nothing on the launch path reads the environment):
 *   VCT_DEBUG_MIP_DENSE          1 = every vct_mipmap reads and writes every tile (the dense build a first frame or an upload pays)
 *   VCT_DEBUG_CONE_VARIANT       -1 = automatic, 0 = literal shader loop, 1 = every fetch blends two levels, 2 = one warp per cone slot,
 *                                3 = all diffuse cones of a tile in one warp
 *   VCT_DEBUG_CONE_GRID          1 = cone kernel on a grid sized by the host instead of the persistent work queue
 *   VCT_DEBUG_CONE_RESERVE_SMS   k = CTAs of the persistent cone kernel that land on the last k SMs retire at once (experiment: SMs left to
 *                                another frame's front half; profiles/r02_frames_in_flight.txt)
 *   VCT_DEBUG_TRACE_LOW_PRIORITY 1 = FRAMES IN FLIGHT: cones + shade of this device object run on the one lowest-priority stream that all
 *                                device objects of the GPU share.  Two device objects (each with its scene / grid / target) rendering
 *                                alternate frames then overlap the front half of frame i+1 with the trace of frame i; results are
 *                                identical to the plain loop
 *   VCT_DEBUG_PEER_REPLICATE     multi-GPU voxelization, set before vct_peer_connect: 1 = every rank voxelizes the whole scene (no voxel
 *                                exchange), 0 = z-slabs + sparse voxel push, -1 = by scene size (<= 16 k triangles: replicate)
 *   VCT_DEBUG_CONE_CTAS_PER_SM   n = CTAs per SM of the persistent cone kernel (experiment; 0 = all that fit)
 *   VCT_DEBUG_SMALL_LIMIT        p = bounding-box size in pixels up to which the set-up kernels rasterise a triangle by its own lane
 *                                (-1 = built-in 100; tools/small_limit_sweep.py) */
#define VCT_DEBUG_MIP_DENSE 1
#define VCT_DEBUG_CONE_VARIANT 2
#define VCT_DEBUG_CONE_GRID 3
#define VCT_DEBUG_CONE_RESERVE_SMS 4
#define VCT_DEBUG_TRACE_LOW_PRIORITY 5
#define VCT_DEBUG_PEER_REPLICATE 6
#define VCT_DEBUG_CONE_CTAS_PER_SM 7
#define VCT_DEBUG_SMALL_LIMIT 8
int vct_debug_set(vct_device_t* dev, int key, int value);

/* ---- multi-GPU (no reference counterpart: the reference is single-GPU).  One process per GPU on one node; the exchange
 *      runs over NVLink peer memory (CUDA IPC), fused into the producing kernels -- see csrc/peer.cu.  Protocol:
 *        every rank:  vct_peer_export(...)  ->  exchange the handles (any transport: torch.distributed, MPI, a file)
 *                     vct_peer_connect(..., all_handles, frame_root)  ->  process barrier  ->  vct_render_frame per frame
 *      After the connect, vct_render_frame on that grid/target renders this rank's share.  Scenes of more than 16 k triangles: it voxelizes
 *      its z-slab [rank*R/n, (rank+1)*R/n) (integer division), stores the resolved voxels into every peer's grid and builds the mip chain
 *      locally once all slabs have arrived.  Smaller scenes (voxelization is launch latency, not work): every rank voxelizes the whole scene
 *      into its own grid and no voxel crosses NVLink (VCT_DEBUG_PEER_REPLICATE forces either mode; decided at the first frame of a
 *      connection, from a scene that must be the same on every rank).  Either way the rank rasterises the G-buffer of and traces the
 *      32x32 screen tiles (tx, ty) with (tx + k ty) % n == rank (vct_trace_params_t.tile_rank) and stores the finished pixels into the
 *      frame of rank `frame_root` (-1: of every rank).  All ranks must call vct_render_frame the same number of times.
 *      The grid and the frame are bit-identical to the single-GPU result.  A peer that never signals (dead, out of step) makes the
 *      next call on the device return VCT_ERR_CUDA after ~5 s. ---- */
typedef struct { unsigned char bytes[320]; } vct_peer_handle_t;
int vct_peer_export(vct_device_t* dev, vct_grid_t* g, vct_target_t* t, vct_peer_handle_t* out);
int vct_peer_connect(vct_device_t* dev, vct_grid_t* g, vct_target_t* t, int rank, int nranks, const vct_peer_handle_t* all_ranks, int frame_root);
int vct_peer_disconnect(vct_device_t* dev);
int vct_peer_error(vct_device_t* dev); /* VCT_ERR_CUDA if a flag wait timed out (~5 s) since the connect; synchronises */

/* ---- texture_3d.h:6-10, one generic RGBA8 3-D texture with a mip chain ---- */
int vct_tex3d_create(vct_device_t* dev, int width, int height, int depth, int levels, vct_tex3d_t** out); /* create_tex_3d */
int vct_tex3d_destroy(vct_tex3d_t* t);                                                                     /* destroy_tex_3d */
int vct_tex3d_clear(vct_tex3d_t* t, const float clear_color[4]);                                           /* clear_tex_3d (level 0) */
int vct_tex3d_mip(vct_tex3d_t* t);                                                                          /* mip_tex_3d (glGenerateMipmap, 2x2x2 box) */
int vct_tex3d_upload(vct_tex3d_t* t, int level, const uint32_t* host);
int vct_tex3d_download(vct_tex3d_t* t, int level, uint32_t* host);

#ifdef __cplusplus
}
#endif
#endif /* VCT_C_H */
