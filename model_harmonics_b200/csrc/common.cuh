// Shared helpers for the mhb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdint>

namespace mhb {

// thread-local error text returned by mhb200_last_error()
char* err_buf();
void set_err(const char* fmt, ...);
extern std::atomic<unsigned long long> g_launches;
void profile_begin(cudaStream_t st);
void profile_end(cudaStream_t st);

#define MHB_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess) {                                                        \
      ::mhb::set_err("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return MHB200_ECUDA;                                                          \
    }                                                                               \
  } while (0)

// fp64 tensor-core tile: D(8x8) += A(8x4, row) * B(4x8, col).  On sm_100a this is one
// DMMA.8x8x4 and runs at the full fp64 pipe rate (37.0 TFLOP/s measured); the wider
// m16n8k8 shape lowers to four DMMA.8x8x4 at HALF the rate (18.5 TFLOP/s measured,
// profiles/r01_fp64_peaks.jsonl), so only this shape is used.
// fragment ownership (lane = threadIdx.x % 32):
//   a = A[lane/4][lane%4]   b = B[lane%4][lane/4]   c0,c1 = C[lane/4][2*(lane%4) + {0,1}]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// integer power by repeated squaring (exponent >= 0)
__device__ __forceinline__ double ipow(double b, int e) {
  double r = 1.0;
  while (e) {
    if (e & 1) r *= b;
    b *= b;
    e >>= 1;
  }
  return r;
}

__device__ __forceinline__ double ld_as_double(const double* p) { return __ldg(p); }
__device__ __forceinline__ double ld_as_double(const float* p) { return (double)__ldg(p); }

}  // namespace mhb
