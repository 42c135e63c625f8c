// Microbenchmark: fp64 DFMA vs DMMA (mma.sync m8n8k4 / m16n8k8 f64) peak on sm_100a,
// and whether the two pipes overlap when issued from different warps of the same SM.
// Output: one JSON line per experiment.  Used to define the FP64 roofline denominator
// (MEASURED_PEAKS.json has none) and to choose tensor vs CUDA-core fp64 per stage.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

// mode bit0: DFMA warps active, bit1: DMMA warps active.  Warps [0, nw_fma) do DFMA, the rest DMMA.
template <int NACC>
__global__ void __launch_bounds__(1024) k_mix(double *out, int iters, int nw_fma, int variant) {
  int warp = threadIdx.x >> 5;
  double seed = out[0];              // 0.0 at run time; defeats constant folding
  if (warp < nw_fma) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = seed + i;
    double a = 1.0 + seed, b = seed * 0.5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    if (s == 123.456) out[1] = s;
  } else {
    if (variant == 0) {
      double c[NACC][2];
#pragma unroll
      for (int i = 0; i < NACC; ++i) { c[i][0] = seed; c[i][1] = seed; }
      double a = seed + threadIdx.x, b = seed;
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
      }
      double s = 0;
#pragma unroll
      for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
      if (s == 123.456) out[2] = s;
    } else {
      double c[NACC / 2][4];
#pragma unroll
      for (int i = 0; i < NACC / 2; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = seed; }
      double a[4] = {seed, seed, seed, seed}, b[2] = {seed, seed};
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC / 2; ++i) dmma1688(c[i], a, b);
      }
      double s = 0;
#pragma unroll
      for (int i = 0; i < NACC / 2; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
      if (s == 123.456) out[2] = s;
    }
  }
}

// dependent-chain latency: one warp, one chain
__global__ void k_lat(double *out, long long *cyc, int iters) {
  double seed = out[0];
  double a = 1.0 + seed, b = seed, x = seed + 1.0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) { x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); }
  long long t1 = clock64();
  double c0 = seed, c1 = seed, fa = seed, fb = seed;
  long long t2 = clock64();
  for (int it = 0; it < iters; ++it) { dmma884(c0, c1, fa, fb); dmma884(c0, c1, fa, fb); dmma884(c0, c1, fa, fb); dmma884(c0, c1, fa, fb); }
  long long t3 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t3 - t2; }
  if (x + c0 + c1 == 123.456) out[3] = x;
}

static float run(int grid, int block, int iters, int nw_fma, int variant, double *d) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_mix<8><<<grid, block>>>(d, iters, nw_fma, variant);   // warm-up
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    CK(cudaEventRecord(e0));
    k_mix<8><<<grid, block>>>(d, iters, nw_fma, variant);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double *d; CK(cudaMalloc(&d, 64)); CK(cudaMemset(d, 0, 64));
  long long *cyc; CK(cudaMalloc(&cyc, 16));
  const int iters = 20000, NACC = 8;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
  for (int ctas = 1; ctas <= 2; ++ctas)
  for (int nw = 4; nw <= 32; nw *= 2) {
    if (ctas * nw > 64) continue;
    int block = nw * 32, grid = sms * ctas;
    // pure DFMA
    float ms = run(grid, block, iters, nw, 0, d);
    double fl = 2.0 * grid * (double)block * iters * NACC;
    printf("{\"exp\": \"dfma\", \"warps_per_cta\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.2f}\n", nw, ctas, ms, fl / ms * 1e-9);
    for (int variant = 0; variant < 2; ++variant) {
      ms = run(grid, block, iters, 0, variant, d);
      double flm = 2.0 * grid * (double)nw * iters * NACC * 256.0;
      printf("{\"exp\": \"%s\", \"warps_per_cta\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.2f}\n",
             variant ? "dmma_m16n8k8" : "dmma_m8n8k4", nw, ctas, ms, flm / ms * 1e-9);
    }
    // mixed: half the warps DFMA, half DMMA (m8n8k4).  If pipes are separate, time ~ max of halves.
    ms = run(grid, block, iters, nw / 2, 0, d);
    double f1 = 2.0 * grid * (double)(nw / 2) * 32 * iters * NACC;
    double f2 = 2.0 * grid * (double)(nw / 2) * iters * NACC * 256.0;
    printf("{\"exp\": \"mixed_half_half\", \"warps_per_cta\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"dfma_tflops\": %.2f, \"dmma_tflops\": %.2f, \"sum_tflops\": %.2f}\n",
           nw, ctas, ms, f1 / ms * 1e-9, f2 / ms * 1e-9, (f1 + f2) / ms * 1e-9);
  }
  // mixed with time-balanced split: DMMA does 8x the flops per warp-instruction, so 8 DFMA warps : 1 DMMA warp ~ equal pipe time if same rate
  {
    int nw = 18, block = nw * 32, grid = sms;
    float ms = run(grid, block, iters, 16, 0, d);
    double f1 = 2.0 * grid * 16.0 * 32 * iters * NACC, f2 = 2.0 * grid * 2.0 * iters * NACC * 256.0;
    printf("{\"exp\": \"mixed_16fma_2mma\", \"ms\": %.4f, \"dfma_tflops\": %.2f, \"dmma_tflops\": %.2f, \"sum_tflops\": %.2f}\n", ms, f1 / ms * 1e-9, f2 / ms * 1e-9, (f1 + f2) / ms * 1e-9);
    ms = run(grid, block, iters, 16, 0, d);
  }
  k_lat<<<1, 32>>>(d, cyc, 1000); CK(cudaDeviceSynchronize());
  long long h[2]; CK(cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost));
  printf("{\"exp\": \"latency\", \"dfma_cycles\": %.2f, \"dmma884_cycles\": %.2f}\n", h[0] / 4000.0, h[1] / 4000.0);
  // sustained run (~3 s) of the best DFMA config to see clocks under fp64 load
  {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    int reps = 60;
    for (int r = 0; r < reps; ++r) k_mix<8><<<sms * 2, 512>>>(d, 200000, 16, 0);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double fl = 2.0 * sms * 2 * 512.0 * 200000 * NACC * reps;
    printf("{\"exp\": \"dfma_sustained\", \"ms\": %.1f, \"tflops\": %.2f}\n", ms, fl / ms * 1e-9);
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) k_mix<8><<<sms * 2, 512>>>(d, 25000, 0, 0);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    fl = 2.0 * sms * 2 * 16.0 * 25000 * NACC * 256.0 * reps;
    printf("{\"exp\": \"dmma_sustained\", \"ms\": %.1f, \"tflops\": %.2f}\n", ms, fl / ms * 1e-9);
  }
  return 0;
}
