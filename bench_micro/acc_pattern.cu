// Does the accumulate_level instruction pattern (stride-8 FMAs into 32 distinct accumulators, powers
// recomputed per level, no memory traffic) reach the DFMA peak at 8 warps/SM?  Compared with 16 and 32 warps.
#include <cstdio>
#include <cuda_runtime.h>
template <bool UPPER>
__device__ __forceinline__ void accumulate_level(double (&acc)[32], double u, double r) {
  const double r2 = r * r, r3 = r2 * r, r4 = r2 * r2;
  const double r5 = r4 * r, r6 = r4 * r2, r7 = r4 * r3, r8 = r4 * r4;
  const double r16 = r8 * r8;
  double us[4];
  us[0] = UPPER ? u * (r16 * r16) : u;
  us[1] = us[0] * r8; us[2] = us[0] * r16; us[3] = us[1] * r16;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const double v = us[s];
    acc[8*s+0] += v; acc[8*s+1] = fma(v, r, acc[8*s+1]); acc[8*s+2] = fma(v, r2, acc[8*s+2]); acc[8*s+3] = fma(v, r3, acc[8*s+3]);
    acc[8*s+4] = fma(v, r4, acc[8*s+4]); acc[8*s+5] = fma(v, r5, acc[8*s+5]); acc[8*s+6] = fma(v, r6, acc[8*s+6]); acc[8*s+7] = fma(v, r7, acc[8*s+7]);
  }
}
__global__ void __launch_bounds__(256, 2) k(double* out, const double2* in, int levels, int reps) {
  extern __shared__ double2 stg[];
  for (int i = threadIdx.x; i < levels * 64; i += blockDim.x) stg[i] = in[i];
  __syncthreads();
  double acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.0;
  const int col = threadIdx.x & 63;
  const bool upper = (threadIdx.x >> 6) & 1;
  for (int rp = 0; rp < reps; ++rp) {
    if (upper) {
#pragma unroll 2
      for (int kk = 0; kk < levels; ++kk) { const double2 e = stg[kk * 64 + col]; accumulate_level<true>(acc, e.x, e.y); }
    } else {
#pragma unroll 2
      for (int kk = 0; kk < levels; ++kk) { const double2 e = stg[kk * 64 + col]; accumulate_level<false>(acc, e.x, e.y); }
    }
  }
  double s = 0; for (int j = 0; j < 32; ++j) s += acc[j];
  if (s == 123.456) out[0] = s;
}
int main() {
  const int levels = 37, reps = 400;
  double2* in; cudaMalloc(&in, levels * 64 * sizeof(double2));
  double2 h[37 * 64]; for (int i = 0; i < levels * 64; ++i) h[i] = make_double2(1e-3 * (i % 7), 1.0 + 1e-4 * (i % 13));
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  double* out; cudaMalloc(&out, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int threads = 128; threads <= 256; threads *= 2) for (int ctas = 1; ctas <= 4; ++ctas) {
    if (threads * ctas > 1024 && threads * ctas * 100 > 65536) {}   // register-limited configs simply run fewer CTAs
    k<<<148 * ctas, threads, levels * 64 * sizeof(double2)>>>(out, in, levels, reps);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<<<148 * ctas, threads, levels * 64 * sizeof(double2)>>>(out, in, levels, reps);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // fp64-pipe instructions per thread-level: lower 43, upper 45 -> mean 44; count each as 2 flops (FMA-equivalent)
    const double slots = 148.0 * ctas * threads * (double)levels * reps * 44.0;
    printf("{\"threads\": %d, \"ctas_per_sm\": %d, \"ms\": %.3f, \"pipe_tflops_equiv\": %.2f, \"err\": \"%s\"}\n", threads, ctas, ms,
           2.0 * slots / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
