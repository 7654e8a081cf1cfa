// device.h -- host-side Device of the B200-native build.
//
// Replaces src/device.h:14-38 of the reference: there `class Device` is an empty stub holding a vector of
// `pipeline_state_t{vao, shader_program, textures[32], ubos[32]}` and is not even compiled
// (CMakeLists.txt:5-7).  With OpenGL gone the "pipeline state" is a CUDA device + one stream + the
// scratch arenas of the kernels, all owned by the opaque vct_device_t of the C ABI (vct_c.h).  The
// class keeps the reference's name and its MAX_BOUND_* constants for source compatibility; every
// method is a thin forwarder, no compute and no CPU fallback live here.
#pragma once

#include <cstddef>
#include <cstdint>

#include "vct/vct_c.h"

#define MAX_BOUND_TEXTURES 32
#define MAX_BOUND_UBOS 32

namespace vct {

typedef size_t pipeline_id_t;
typedef size_t vertex_buffer_t;

class Device {
 public:
  // cuda_ordinal: which GPU this Device (and the Renderer that owns it) drives; one Device per rank
  explicit Device(int cuda_ordinal = 0);
  ~Device();
  Device(const Device&) = delete;
  Device& operator=(const Device&) = delete;

  bool ok() const { return m_dev != nullptr; }      // false: no sm_100 GPU / CUDA error; see error()
  const char* error() const;                        // vct_last_error() of the failing call
  void sync();                                      // glFinish() equivalent
  void* stream() const;                             // cudaStream_t (interop with NCCL / torch)
  bool last_frame_timings(float out_ms[8]) const;   // per-stage CUDA-event times of the last Renderer::render()
  vct_device_t* handle() const { return m_dev; }

 private:
  vct_device_t* m_dev = nullptr;
};

}  // namespace vct
