// device.cpp -- vct::Device: owner of the CUDA device handle of the C ABI (see include/vct/device.h).
#include "vct/device.h"

#include <cstdio>

namespace vct {

Device::Device(int cuda_ordinal) {
  if (vct_device_create(cuda_ordinal, &m_dev) != VCT_OK) {
    // the reference prints and carries on (gl_helpers.c:69-81); there is no CPU path to fall back to
    std::fprintf(stderr, "vct::Device: %s\n", vct_last_error());
    m_dev = nullptr;
  }
}

Device::~Device() {
  if (m_dev) vct_device_destroy(m_dev);
}

const char* Device::error() const { return vct_last_error(); }
void Device::sync() { if (m_dev) vct_device_sync(m_dev); }
void* Device::stream() const { return m_dev ? vct_device_stream(m_dev) : nullptr; }
bool Device::last_frame_timings(float out_ms[8]) const { return m_dev && vct_last_frame_timings(m_dev, out_ms) == VCT_OK; }

}  // namespace vct
