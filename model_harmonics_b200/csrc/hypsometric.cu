// Upstream producer of the atmosphere operator's inputs: hypsometric integration of temperature
// and specific humidity on hybrid model levels into geopotential at the levels and pressure
// differences across them (reference model_harmonics/reanalysis/reanalysis_geopotential_heights.py:349-383).
//
// Streaming, four (lat, lon) columns per thread walking the levels bottom-up (the geopotential is a
// running sum), 16-byte loads and stores coalesced along longitude, the next level's loads in flight.
// The reference's mixed precision is reproduced operation by operation with round-to-nearest
// intrinsics (no FMA contraction): float32 T and q widened to float64 for the virtual temperature,
// float64 level pressures and logarithm, float32 running geopotential updated as
// float32(float64(gh) + term), float32 outputs.  The only operation that is not bit-defined is
// the logarithm of the pressure ratio (log_ratio below: within an ulp of glibc's log of the rounded
// quotient), which very rarely moves a float32 rounding of the running sum; the pressure differences
// are bit-identical.
#include "../../include/mhb200.h"
#include "common.cuh"

#include <algorithm>
#include <cmath>

namespace mhb {

constexpr int HYP_MAX_COEF = 160;      // hybrid coefficients travel as kernel parameters

struct HypParams {
  const float *T, *Q, *SP, *gh0;
  float *Z, *DIFF;
  long long ncol;
  int nlev, nA, nAI;
  double R_dry, Pmin;
  float divisor;                       // outputs Z / divisor (float32 division), 1 = geopotential as is
  float rdivisor;                      // float32(1 / divisor), correctly rounded
  double A[HYP_MAX_COEF], B[HYP_MAX_COEF], AI[HYP_MAX_COEF], BI[HYP_MAX_COEF];
  double C[HYP_MAX_COEF], LC[HYP_MAX_COEF];      // per level: estimate of Pnum / Pdom and its logarithm (log_ratio)
  double poly[9];                                // 1/3, 1/5, ..., 1/19
};

// log(a / b) for a, b > 0, given a level-wide estimate c of the ratio and lc = log(c).
// The quotient and the general logarithm (~50 fp64-pipe slots per element, which made this stream compute
// bound at 0.3 ms) are replaced by
//   u = (a - c b) / (a + c b),   log(a / b) = lc + 2 atanh(u) = lc + 2 u (1 + u^2/3 + u^4/5 + ...)
// for |u| <= 0.1 (nine terms: truncation < 1e-18) with a Newton division.  The ratio of two level pressures
// depends only weakly on the surface pressure, so with c taken at a reference surface pressure u stays at the
// per-cent level; anything else (|u| > 0.1) takes log(a / b) as the reference writes it.  Both forms are
// within an ulp of the exact value, which is all the float32 running sum can see.
// (the series coefficients 1/3 ... 1/19 travel as kernel parameters so that the FMAs read them straight from
// the constant bank; `u_out` lets the caller test |u| <= 0.1 once per level instead of branching per element)
__device__ __forceinline__ double log_ratio_series(double a, double b, double c, double lc, const double* pc,
                                                   double& u_out) {
  const double cb = c * b;
  const double num = a - cb, den = a + cb;
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(den));
  double e = fma(-den, y, 1.0);
  y = fma(y, e, y);
  double u = num * y;
  e = fma(-den, u, num);
  u = fma(e, y, u);
  u_out = u;
  const double u2 = u * u;
  double pl = pc[8];
#pragma unroll
  for (int i = 7; i >= 0; --i) pl = fma(pl, u2, pc[i]);
  const double t = u + u;
  return lc + fma(t * u2, pl, t);
}

// V columns per thread (V = 4: 16-byte loads and stores), the next level's loads in flight.  The level
// pressures A[k] + B[k] sp are formed once per level and reused as the next level's numerator (the
// reference recomputes them: same expression, same bits).
template <int V>
__global__ void __launch_bounds__(256) hypsometric_kernel(const __grid_constant__ HypParams q) {
  const long long col = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * V;
  if (col >= q.ncol) return;
  struct alignas(4 * V) Vec {
    float v[V];
  };
  auto ld = [&](const float* p) { return *reinterpret_cast<const Vec*>(p); };
  const size_t stride = (size_t)q.ncol;
  const Vec spv = ld(q.SP + col), g0 = ld(q.gh0 + col);
  double sp[V], pk[V], pik[V];
  float gh[V];
#pragma unroll
  for (int c = 0; c < V; ++c) {
    sp[c] = (double)spv.v[c];
    gh[c] = g0.v[c];
    pk[c] = __dadd_rn(q.A[0], __dmul_rn(q.B[0], sp[c]));          // :362, level 0
    pik[c] = __dadd_rn(q.AI[0], __dmul_rn(q.BI[0], sp[c]));       // :375
  }
  const float* Tp = q.T + col;
  const float* Qp = q.Q + col;
  float* Zp = q.Z + col;
  float* Dp = q.DIFF + col;
  const bool scale = (q.divisor != 1.0f);
  Vec tn = ld(Tp), qn = ld(Qp);
  for (int k = 0; k < q.nlev; ++k) {
    const Vec tv = tn, qv = qn;
    if (k + 1 < q.nlev) {
      tn = ld(Tp + (size_t)(k + 1) * stride);
      qn = ld(Qp + (size_t)(k + 1) * stride);
    }
    // past the last coefficient the denominators are the Pmin floor (:363-368, :377-382): with the pair
    // (Pmin, 0) the same expression max(a + b sp, Pmin) yields it, with no per-element select
    const bool lastA = (k + 1 == q.nA), lastI = (k + 1 == q.nAI);
    const double a1 = lastA ? q.Pmin : q.A[k + 1], b1 = lastA ? 0.0 : q.B[k + 1];
    const double ai1 = lastI ? q.Pmin : q.AI[k + 1], bi1 = lastI ? 0.0 : q.BI[k + 1];
    const double ck = q.C[k], lck = q.LC[k];
    double lg[V], Pdom[V], umax = 0.0;
#pragma unroll
    for (int c = 0; c < V; ++c) {
      const double pnext = __dadd_rn(a1, __dmul_rn(b1, sp[c]));        // :362-368
      Pdom[c] = fmax(pnext, q.Pmin);
      double u;
      lg[c] = log_ratio_series(pk[c], Pdom[c], ck, lck, q.poly, u);
      umax = fmax(umax, fabs(u));
      if (u != u) umax = 1.0;
      pk[c] = pnext;
    }
    if (umax > 0.1) {
      // ratio far from the level's estimate: the logarithm as the reference writes it
      // (pk was advanced above: the numerator of this level is recomputed)
#pragma unroll
      for (int c = 0; c < V; ++c) {
        const double pnum = __dadd_rn(q.A[k], __dmul_rn(q.B[k], sp[c]));
        lg[c] = log(__ddiv_rn(pnum, Pdom[c]));
      }
    }
    Vec zo, dfo;
#pragma unroll
    for (int c = 0; c < V; ++c) {
      // :360  Tv = (1.0 + 0.609133 * QV) * T   (float64 on the widened values)
      const double Tv = __dmul_rn(__dadd_rn(1.0, __dmul_rn(0.609133, (double)qv.v[c])), (double)tv.v[c]);
      // :371  gh += R_dry * Tv * log(Pnum / Pdom)
      const double term = __dmul_rn(__dmul_rn(q.R_dry, Tv), lg[c]);
      gh[c] = __double2float_rn(__dadd_rn((double)gh[c], term));
      // :373 (and the float32 division by GRAVITY of reanalysis_atmospheric_harmonics.py:357): correctly
      // rounded quotient by a constant from its correctly rounded reciprocal (Markstein): q0 = gh r,
      // q = q0 + (gh - q0 d) r with the residual exact in the FMA
      float zq = gh[c];
      if (scale) {
        const float q0 = gh[c] * q.rdivisor;
        zq = fmaf(fmaf(-q0, q.divisor, gh[c]), q.rdivisor, q0);
      }
      zo.v[c] = zq;
      // :375-383
      const double pinext = __dadd_rn(ai1, __dmul_rn(bi1, sp[c]));
      dfo.v[c] = __double2float_rn(__dsub_rn(fmax(pinext, q.Pmin), pik[c]));
      pik[c] = pinext;
    }
    *reinterpret_cast<Vec*>(Zp + (size_t)k * stride) = zo;
    *reinterpret_cast<Vec*>(Dp + (size_t)k * stride) = dfo;
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
  q.rdivisor = 1.0f / divisor;
  for (int i = 0; i < HYP_MAX_COEF; ++i) {
    q.A[i] = (i < nA) ? A[i] : 0.0;
    q.B[i] = (i < nA) ? B[i] : 0.0;
    q.AI[i] = (i < nAI) ? AI[i] : 0.0;
    q.BI[i] = (i < nAI) ? BI[i] : 0.0;
  }
  for (int i = 0; i < 9; ++i) q.poly[i] = 1.0 / (double)(2 * i + 3);
  // level-wide estimate of the pressure ratio Pnum / Pdom (at a reference surface pressure) and its logarithm
  for (int k = 0; k < HYP_MAX_COEF; ++k) {
    double c = 1.0;
    if (k < nlev) {
      const double sp_ref = 1.0e5;
      const double pn = A[k] + B[k] * sp_ref;
      const double pd = (k + 1 == nA) ? Pmin : std::max(A[k + 1] + B[k + 1] * sp_ref, Pmin);
      if (pn > 0.0 && pd > 0.0) c = pn / pd;
    }
    q.C[k] = c;
    q.LC[k] = log(c);
  }
  // four columns per thread when every row of every array starts on a 16-byte boundary
  const bool vec = (ncol % 4 == 0) && ((((uintptr_t)T | (uintptr_t)Q | (uintptr_t)SP | (uintptr_t)gh0 |
                                        (uintptr_t)Z_out | (uintptr_t)DIFF_out) & 15) == 0);
  profile_begin((cudaStream_t)stream);
  if (vec) {
    const unsigned grid = (unsigned)((ncol / 4 + 255) / 256);
    hypsometric_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(q);
  } else {
    const unsigned grid = (unsigned)((ncol + 255) / 256);
    hypsometric_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(q);
  }
  profile_end((cudaStream_t)stream);
  g_launches.fetch_add(1);
  MHB_CUDA(cudaGetLastError());
  return MHB200_OK;
}
