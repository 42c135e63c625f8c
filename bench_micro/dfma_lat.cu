// Dependent-issue latency of DFMA / DMUL on sm_100a: one warp per SM sub-partition, NCH independent chains.
#include <cstdio>
#include <cuda_runtime.h>
template <int NCH>
__global__ void k(double* out, int iters, long long* cyc) {
  double a[NCH];
  const double m = 1.0 + out[0], b = out[0] * 0.5;
#pragma unroll
  for (int i = 0; i < NCH; ++i) a[i] = out[0] + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < NCH; ++i) a[i] = fma(a[i], m, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) s += a[i];
  if (s == 123.456) out[1] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int NCH>
void run(double* out, long long* cyc, int warps) {
  const int iters = 2000;
  k<NCH><<<1, 32 * warps>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("{\"chains\": %d, \"warps_in_cta\": %d, \"cycles_per_dfma_per_warp\": %.2f}\n", NCH, warps,
         (double)h / (iters * 8.0 * NCH));
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 64); cudaMemset(out, 0, 64); cudaMalloc(&cyc, 8);
  for (int w : {1, 4, 8, 16}) {
    run<1>(out, cyc, w); run<2>(out, cyc, w); run<4>(out, cyc, w); run<8>(out, cyc, w); run<16>(out, cyc, w);
  }
  return 0;
}
