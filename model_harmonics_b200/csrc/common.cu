// Library-wide state: error text, launch counter, device probe.
#include "../../include/mhb200.h"
#include "common.cuh"

namespace mhb {

std::atomic<unsigned long long> g_launches{0};

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

void set_err(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
}

}  // namespace mhb

extern "C" {

int mhb200_version(void) { return 100; }

const char* mhb200_last_error(void) { return mhb::err_buf(); }

uint64_t mhb200_launch_count(void) { return (uint64_t)mhb::g_launches.load(); }

int mhb200_device_check(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    mhb::set_err("no CUDA device available (%s); this library has no CPU fallback",
                 e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return MHB200_ENODEVICE;
  }
  if (device < 0 || device >= n) {
    mhb::set_err("device %d out of range (0..%d)", device, n - 1);
    return MHB200_EINVAL;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    mhb::set_err("cudaGetDeviceProperties failed: %s", cudaGetErrorString(e));
    return MHB200_ECUDA;
  }
  if (prop.major != 10) {
    mhb::set_err("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    return MHB200_ENODEVICE;
  }
  return MHB200_OK;
}

}  // extern "C"
