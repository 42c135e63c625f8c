// Probe of the TMA path used by atm_moment_kernel: 4-D tiled map over a (field, level, lat, lon) cube,
// box = 64 lon x 1 lat x 4 levels, loaded at a given start longitude.  usage: tma_probe <f64|f32> <x0> [swz]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <typename T>
__global__ void probe(const __grid_constant__ CUtensorMap tm, int x0, int h, int lev, T* out) {
  extern __shared__ __align__(1024) unsigned char sm[];
  T* buf = reinterpret_cast<T*>(sm);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 4 * 64 * sizeof(T));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"((uint32_t)(4 * 64 * sizeof(T)))
                 : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(buf)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(bar)), "r"(x0), "r"(h), "r"(lev), "r"(0)
        : "memory");
  }
  asm volatile(
      "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(
          smem_u32(bar)),
      "r"(0)
      : "memory");
  for (int i = threadIdx.x; i < 4 * 64; i += blockDim.x) out[i] = buf[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename T>
int run(int x0, CUtensorMapDataType dt) {
  const int nlon = 1440, nlat = 8, nlev = 6;
  std::vector<T> host((size_t)nlon * nlat * nlev);
  for (size_t i = 0; i < host.size(); ++i) host[i] = (T)(i % 100003);
  T *dev, *out;
  cudaMalloc(&dev, host.size() * sizeof(T));
  cudaMalloc(&out, 256 * sizeof(T));
  cudaMemcpy(dev, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice);
  void* f = nullptr;
  cudaDriverEntryPointQueryResult qr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) != cudaSuccess || !f) {
    printf("no encoder\n");
    return 2;
  }
  CUtensorMap tm;
  cuuint64_t dims[4] = {(cuuint64_t)nlon, (cuuint64_t)nlat, (cuuint64_t)nlev, 1};
  cuuint64_t strides[3] = {nlon * sizeof(T), (cuuint64_t)nlon * nlat * sizeof(T), (cuuint64_t)nlon * nlat * nlev * sizeof(T)};
  cuuint32_t box[4] = {64, 1, 4, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = ((EncodeTiledFn)f)(&tm, dt, 4, dev, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc %d\n", (int)r);
  if (r != CUDA_SUCCESS) return 3;
  const int h = 3, lev = 1;
  probe<T><<<1, 64, 4 * 64 * sizeof(T) + 64>>>(tm, x0, h, lev, out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 4;
  std::vector<T> got(256);
  cudaMemcpy(got.data(), out, 256 * sizeof(T), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int k = 0; k < 4; ++k)
    for (int i = 0; i < 64; ++i) {
      const int col = x0 + i;
      T want = 0;
      if (col >= 0 && col < nlon) want = host[((size_t)(lev + k) * nlat + h) * nlon + col];
      if (got[k * 64 + i] != want) ++bad;
    }
  printf("mismatches %d\n", bad);
  return bad ? 5 : 0;
}

int main(int argc, char** argv) {
  const bool f64 = argc > 1 && !strcmp(argv[1], "f64");
  const int x0 = argc > 2 ? atoi(argv[2]) : 0;
  return f64 ? run<double>(x0, CU_TENSOR_MAP_DATA_TYPE_FLOAT64) : run<float>(x0, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
}
