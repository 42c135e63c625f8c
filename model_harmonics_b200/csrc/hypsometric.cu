// Upstream producer of the atmosphere operator's inputs: hypsometric integration of temperature
// and specific humidity on hybrid model levels into geopotential at the levels and pressure
// differences across them (reference model_harmonics/reanalysis/reanalysis_geopotential_heights.py:349-383).
//
// Streaming, one thread per (lat, lon) column walking the levels bottom-up (the geopotential is a
// running sum), loads and stores coalesced along longitude, four levels of loads in flight.
// The reference's mixed precision is reproduced operation by operation with round-to-nearest
// intrinsics (no FMA contraction): float32 T and q widened to float64 for the virtual temperature,
// float64 level pressures and logarithm, float32 running geopotential updated as
// float32(float64(gh) + term), float32 outputs.  The only operation that is not bit-defined is
// log(): CUDA's is within 1 ulp of glibc's, which very rarely moves a float32 rounding.
#include "../../include/mhb200.h"
#include "common.cuh"

namespace mhb {

constexpr int HYP_MAX_COEF = 160;      // hybrid coefficients travel as kernel parameters

struct HypParams {
  const float *T, *Q, *SP, *gh0;
  float *Z, *DIFF;
  long long ncol;
  int nlev, nA, nAI;
  double R_dry, Pmin;
  float divisor;                       // outputs Z / divisor (float32 division), 1 = geopotential as is
  double A[HYP_MAX_COEF], B[HYP_MAX_COEF], AI[HYP_MAX_COEF], BI[HYP_MAX_COEF];
};

__global__ void __launch_bounds__(256) hypsometric_kernel(const __grid_constant__ HypParams q) {
  const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= q.ncol) return;
  const double sp = (double)q.SP[col];
  float gh = q.gh0[col];
  const float* Tp = q.T + col;
  const float* Qp = q.Q + col;
  float* Zp = q.Z + col;
  float* Dp = q.DIFF + col;
  const size_t stride = (size_t)q.ncol;
  constexpr int U = 4;
  for (int k0 = 0; k0 < q.nlev; k0 += U) {
    float tv[U], qv[U];
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const int k = min(k0 + i, q.nlev - 1);
      tv[i] = __ldg(Tp + (size_t)k * stride);
      qv[i] = __ldg(Qp + (size_t)k * stride);
    }
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const int k = k0 + i;
      if (k < q.nlev) {
        // :360  Tv = (1.0 + 0.609133 * QV) * T   (float64 on the widened values)
        const double Tv = __dmul_rn(__dadd_rn(1.0, __dmul_rn(0.609133, (double)qv[i])), (double)tv[i]);
        // :362-368
        const double Pnum = __dadd_rn(q.A[k], __dmul_rn(q.B[k], sp));
        const double Pdom = (k + 1 == q.nA) ? q.Pmin : fmax(__dadd_rn(q.A[k + 1], __dmul_rn(q.B[k + 1], sp)), q.Pmin);
        // :371  gh += R_dry * Tv * log(Pnum / Pdom)
        const double term = __dmul_rn(__dmul_rn(q.R_dry, Tv), log(__ddiv_rn(Pnum, Pdom)));
        gh = __double2float_rn(__dadd_rn((double)gh, term));
        Zp[(size_t)k * stride] = (q.divisor == 1.0f) ? gh : __fdiv_rn(gh, q.divisor);     // :373
        // :375-383
        const double Plower = __dadd_rn(q.AI[k], __dmul_rn(q.BI[k], sp));
        const double Pupper =
            (k + 1 == q.nAI) ? q.Pmin : fmax(__dadd_rn(q.AI[k + 1], __dmul_rn(q.BI[k + 1], sp)), q.Pmin);
        Dp[(size_t)k * stride] = __double2float_rn(__dsub_rn(Pupper, Plower));
      }
    }
  }
}

}  // namespace mhb

using namespace mhb;

extern "C" int mhb200_geopotential_levels_f32(int device, const float* T, const float* Q, const float* SP,
                                              const float* gh0, const double* A, const double* B, int nA,
                                              const double* AI, const double* BI, int nAI, int nlev,
                                              int64_t ncol, double R_dry, double Pmin, float divisor, float* Z_out,
                                              float* DIFF_out, void* stream) {
  if (!T || !Q || !SP || !gh0 || !A || !B || !AI || !BI || !Z_out || !DIFF_out || nlev < 1 || ncol < 1 ||
      nA < nlev || nAI < nlev || nA > HYP_MAX_COEF || nAI > HYP_MAX_COEF || !(divisor > 0.0f)) {
    set_err("mhb200_geopotential_levels_f32: bad argument (needs nlev <= nA, nAI <= %d)", HYP_MAX_COEF);
    return MHB200_EINVAL;
  }
  int rc = mhb200_device_check(device);
  if (rc != MHB200_OK) return rc;
  MHB_CUDA(cudaSetDevice(device));
  HypParams q;
  q.T = T;
  q.Q = Q;
  q.SP = SP;
  q.gh0 = gh0;
  q.Z = Z_out;
  q.DIFF = DIFF_out;
  q.ncol = ncol;
  q.nlev = nlev;
  q.nA = nA;
  q.nAI = nAI;
  q.R_dry = R_dry;
  q.Pmin = Pmin;
  q.divisor = divisor;
  for (int i = 0; i < HYP_MAX_COEF; ++i) {
    q.A[i] = (i < nA) ? A[i] : 0.0;
    q.B[i] = (i < nA) ? B[i] : 0.0;
    q.AI[i] = (i < nAI) ? AI[i] : 0.0;
    q.BI[i] = (i < nAI) ? BI[i] : 0.0;
  }
  const unsigned grid = (unsigned)((ncol + 255) / 256);
  profile_begin((cudaStream_t)stream);
  hypsometric_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(q);
  profile_end((cudaStream_t)stream);
  g_launches.fetch_add(1);
  MHB_CUDA(cudaGetLastError());
  return MHB200_OK;
}
