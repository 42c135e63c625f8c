// Contracted-moment form of the gridded fields -> Stokes path on B200 (sm_100a).
//
// Replaces the same reference loops as stokes_grid.cu (paths relative to the reference tree):
//   model_harmonics/gen_atmosphere_stokes.py:217-275    geometry pre-pass, level loop, einsum
//   model_harmonics/gen_pressure_stokes.py:194-206      per-degree power pass + complex einsum
//
// Per latitude ring h the radius ratio is written r = r0(h) (1 + d) with a ring-wide expansion
// point r0(h), so that  r^n = r0^n sum_j C(n, j) d^j  (BC05 expands in e = (1 + d)^2 - 1 instead,
// r^n = r0^n sum_j C(n/2, j) e^j, which spares the square root of the geocentric radius)
// and the per-degree radial factor needs only NMOM = 16 moments per column,
//   atmosphere:  M_j[p] = sum_k (dp_k / gamma_k) d_k^j        pressure:  M_j[p] = (P/G) d^j .
// The binomial expansion, the Fourier sum over longitude and the Legendre sum over latitude are all
// linear, so they commute; the expansion is applied LAST, once per coefficient:
//   K_A  (atm_moment_kernel / pressure_moment_kernel)
//        F[slot][m][cs][j] = sum_p {cos, sin}(m phi_p) M_j[p]         DMMA m8n8k4, N = 16 moments
//        (a slot is one ring, or the part of a ring a CTA owned); atmosphere cubes arrive through a
//        TMA pipeline (cp.async.bulk.tensor 4-D boxes: 64 longitudes x 4 levels per fold sub-block)
//   K_B  (legendre_moment_kernel)   per order m, DMMA over slots:
//        G[l][m][cs][j] = sum_slots w_h Plm(x_h) r0(h)^(l+n0) F[slot][m][cs][j]
//        with Plm from the Holmes-Featherstone recursion generated on the fly (never materialised)
//   K_C  (moment_finalize_kernel)
//        C_lm, S_lm = dfactor[l] sum_j C(l + n0, j) G[l][m][{cos, sin}][j]
// Same terms as the reference's sum, different summation order (1e-12 contract, measured ~1e-15).
//
// Validity: the truncated tail is (n |d|)^16 / 16!, so K_A tracks max |d| against 0.8 / n_max and
// raises a flag; the caller then re-runs the call on the direct-power kernels of stokes_grid.cu
// (launched unconditionally, they exit at once while the flag is clear), which are exact for any radius
// like the reference's np.power.
#include "../../include/mhb200.h"
#include "common.cuh"
#include "grid_plan.cuh"

#include <cuda.h>     // CUtensorMap types; the encoder is fetched through cudaGetDriverEntryPoint (no -lcuda)

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace mhb {

constexpr int NMOMC = 16;                 // moments per column
constexpr int MKCF = 64;                  // contraction columns per chunk
constexpr int MKC = 4 * MKCF;             // thread columns per chunk (four fold sub-blocks)
constexpr int MLG = 4;                    // levels per TMA stage
constexpr int MSTAGES = 4;                // stages per CTA (3 measured 5 % slower; 5 do not fit twice per SM)
constexpr int MF_KS = 68;                 // folded-moment buffer: doubles per k-step (2 n-tiles x 32, +4 pad)
constexpr int MF_V = 16 * MF_KS + 8;      // ... per fold variant (+8: variants land on different banks)
constexpr int NB_SUB = 32;                // most fields per K_A launch (see nb_sub: the F workspace stays <= 256 MB)
constexpr int MAX_SPLIT = 8;              // ring splits of K_B
constexpr int G_UNIT = 4 * 8 * 64;        // doubles K_B writes per (field, split, unit): 4 row tiles x 8 n-tiles
constexpr double MOM_CENTER_HEIGHT = 24000.0;   // expansion point above the ellipsoid [m] (atmosphere)
constexpr double MOM_ND_BOUND = 0.8;      // guard: n_max |d| <= 0.8  ->  (0.8)^16 / 16! = 1.3e-15

// ------------------------------------------------------------------------------------------------
// device helpers: mbarrier + TMA (raw PTX), Newton division / square root W elements wide
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// q = n / d and sqrt(a), W independent elements wide (statement by statement across the elements so that
// the instruction stream carries the ILP): fp64 SFU seed, one Newton step, one residual correction --
// correctly rounded in the geometry pass's ranges (bench_micro/newton_check.cu, profiles/r01_newton_check.jsonl)
#ifdef KA_DIV5
template <int W>
__device__ __forceinline__ void div_t(const double (&n)[W], const double (&d)[W], double (&q)[W]) {
  double y[W], e[W];
#pragma unroll
  for (int i = 0; i < W; ++i) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y[i]) : "d"(d[i]));
#pragma unroll
  for (int i = 0; i < W; ++i) e[i] = fma(-d[i], y[i], 1.0);
#pragma unroll
  for (int i = 0; i < W; ++i) y[i] = fma(y[i], e[i], y[i]);
#pragma unroll
  for (int i = 0; i < W; ++i) q[i] = n[i] * y[i];
#pragma unroll
  for (int i = 0; i < W; ++i) e[i] = fma(-d[i], q[i], n[i]);
#pragma unroll
  for (int i = 0; i < W; ++i) q[i] = fma(e[i], y[i], q[i]);
}
#else
// n / d within an ulp in four operations after the seed y (relative error |e| <= 2^-20, e = 1 - d y exact to
// rounding):  n / d = n y (1 + e + e^2 + O(e^3)),  e^3 < 2^-60
template <int W>
__device__ __forceinline__ void div_t(const double (&n)[W], const double (&d)[W], double (&q)[W]) {
  double y[W], e[W];
#pragma unroll
  for (int i = 0; i < W; ++i) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y[i]) : "d"(d[i]));
#pragma unroll
  for (int i = 0; i < W; ++i) e[i] = fma(-d[i], y[i], 1.0);
#pragma unroll
  for (int i = 0; i < W; ++i) q[i] = n[i] * y[i];
#pragma unroll
  for (int i = 0; i < W; ++i) e[i] = fma(e[i], e[i], e[i]);
#pragma unroll
  for (int i = 0; i < W; ++i) q[i] = fma(q[i], e[i], q[i]);
}
#endif
template <int W>
__device__ __forceinline__ void sqrt_t(const double (&a)[W], double (&g)[W]) {
  double h[W], r[W];
#pragma unroll
  for (int i = 0; i < W; ++i) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a[i]));
    g[i] = a[i] * y;
    h[i] = 0.5 * y;
  }
#pragma unroll
  for (int i = 0; i < W; ++i) r[i] = fma(-g[i], h[i], 0.5);
#pragma unroll
  for (int i = 0; i < W; ++i) g[i] = fma(g[i], r[i], g[i]);
#pragma unroll
  for (int i = 0; i < W; ++i) h[i] = fma(h[i], r[i], h[i]);
#pragma unroll
  for (int i = 0; i < W; ++i) r[i] = fma(-g[i], g[i], a[i]);
#pragma unroll
  for (int i = 0; i < W; ++i) g[i] = fma(r[i], h[i], g[i]);
}

// The kernels expand around r0 through a rounded reciprocal c (d = R c - 1, c ~ 1 / (rad_e r0)), so the
// expansion point they really use is 1 / (rad_e c), which differs from the double r0 that K_B raises to the
// n-th power by a relative q ~ 1e-16 -- systematic over a ring and amplified by n.  q is computed here to
// ~1e-32 with error-free products, stored per slot, and K_B multiplies r0^n by (1 - n q).
//   one_plus_q(a, b, c): q with a * b * c = 1 + q;  the squared form: (a * b)^2 * c = 1 + q
__device__ __forceinline__ double residual3(double a, double b, double c) {
  const double ab = a * b, ab_lo = fma(a, b, -ab);
  const double p = ab * c, p_lo = fma(ab, c, -p) + ab_lo * c;
  return (p - 1.0) + p_lo;
}
__device__ __forceinline__ double residual_sq(double a, double b, double c) {
  const double ab = a * b, ab_lo = fma(a, b, -ab);
  const double s = ab * ab, s_lo = fma(ab, ab, -s) + 2.0 * ab * ab_lo;
  const double p = s * c, p_lo = fma(s, c, -p) + s_lo * c;
  return (p - 1.0) + p_lo;
}

// m[j] += w d^j, j < 16:  6 products + 16 accumulates
__device__ __forceinline__ void accumulate16(double (&m)[NMOMC], double w, double d) {
  const double d2 = d * d, d3 = d2 * d, d4 = d2 * d2;
  const double v4 = w * d4, v8 = v4 * d4, v12 = v8 * d4;
  m[0] += w;
  m[1] = fma(w, d, m[1]);
  m[2] = fma(w, d2, m[2]);
  m[3] = fma(w, d3, m[3]);
  m[4] += v4;
  m[5] = fma(v4, d, m[5]);
  m[6] = fma(v4, d2, m[6]);
  m[7] = fma(v4, d3, m[7]);
  m[8] += v8;
  m[9] = fma(v8, d, m[9]);
  m[10] = fma(v8, d2, m[10]);
  m[11] = fma(v8, d3, m[11]);
  m[12] += v12;
  m[13] = fma(v12, d, m[13]);
  m[14] = fma(v12, d2, m[14]);
  m[15] = fma(v12, d3, m[15]);
}

// ------------------------------------------------------------------------------------------------
// K_A, atmosphere
// ------------------------------------------------------------------------------------------------
struct AtmMomParams {
  int nlat, nlon, nlev, nchunks, ngroups;
  int h0, nrings;              // rings [h0, h0 + nrings) of every field of the launch
  int t0;                      // first field of the launch (index into field_scale)
  int fields_in_map;           // 1: all fields share the base cube (coordinate 0)
  const int* cta_unit;         // [ncta + 1] unit span of every CTA; unit = (field, ring, chunk) linearised
  const int* cta_slot;         // [ncta] first slot of the CTA's span
  int slot_base;               // slots of this launch start here
  const int* box_start;        // [nchunks][4] first longitude of the TMA box of every fold sub-block
  const int* col_grid;         // [nchunks][256] grid column of thread column (warp, lane), -1 = inactive
  int dir[4];                  // +1: sub-block runs ascend with the contraction column, -1: descend
  const double* ring_tab;      // caller's 9 x nlat table (see mhb200.h)
  const double* ringc;         // [nlat][RINGC] per-ring constants of the moment form (atm_ring_kernel)
  const double* geoid;
  long long geo_s0, geo_s1;
  const double* field_scale;   // device, 2 per field, or null
  const double* trig;          // A fragments [nchunks * 16][16 row tiles][32 lanes], order block 0
  double* F;                   // [slot][64 orders][cs][16 moments]
  double* slot_r0;
  double* slot_q;              // correction of r0^n per unit of n (see residual3)
  int* slot_ring;
  int* flag;
  double rad_e, inv_rad, hc;
  unsigned dbound_hi;          // high word of the |d| bound
};

struct ColumnC {
  // BC05: a1, a2 orthometric-height factors (field scale folded in), st, ct, g0..g2 normal gravity
  // all lengths scaled by c = 1 / (rad_e r0) = cinv:  H' = z (a1 + a2 z) with a1, a2 scaled, NG = c (N + geoid),
  // st = lam = 2 cos(theta) K and ct = mu = K^2 - 1 with K = -(c N e^2) cos(theta) (see moment_elements),
  // gam = g0 + H' (g1 + g2 H') with g1, g2 rescaled to H' 
  // SW02: rad_e, sg, inv_r0, c0 = geoid / (rad_e r0) - 1
  double a1, a2, st, ct, g0, g1, g2, NG, NG1, cinv;
  double rad_e, sg, inv_r0, c0;
};

// (w, d) of W (level, column) elements:  w = dp / gamma (BC05) or dp (SW02; 1/g0 is applied at the flush),
// d = (r / r0)^2 - 1 (BC05) or r / r0 - 1 (SW02).  Reference lines: gen_atmosphere_stokes.py:222-238 (BC05), :258-260 (SW02).
// X^2 + Y^2 is formed as ((N + geoid + H) sin(theta))^2: cos^2(phi) + sin^2(phi) differs from 1 by rounding only.
template <int MODE, int W>
__device__ __forceinline__ void moment_elements(const double (&z)[W], const double (&dpv)[W], const ColumnC& c,
                                                double (&w)[W], double (&d)[W]) {
  if (MODE == MHB200_METHOD_BC05) {
    // BC05 expands in e = (R / (rad_e r0))^2 - 1 = 2 d + d^2 with the generalised binomial series
    // r^n = r0^n (1 + e)^(n/2) = r0^n sum_j C(n/2, j) e^j  (same convergence: (n/2) e ~ n d) -- R^2 is what
    // the geometry produces, so no square root is needed
    // With c = 1 / (rad_e r0) folded into the column constants (H' = c H, A' = c (N + geoid + H)) and
    // K = -c N e^2 cos(theta):
    //   e = (A' sin)^2 + (A' cos + K)^2 - 1 = A'^2 (sin^2 + cos^2) + A' (2 cos K) + (K^2 - 1) = A' (A' + lam) + mu
    // (5 operations with H').  sin^2 + cos^2 of the table's own values is 1 + O(1e-16), and mu rounds at 1e-16:
    // both residuals are per-ring constants, computed exactly at the start of the unit and folded into the
    // slot's correction q (see the caller), so nothing systematic is left at n * 1e-16.
    double H[W], gam[W];
#pragma unroll
    for (int i = 0; i < W; ++i) H[i] = z[i] * fma(c.a2, z[i], c.a1);
#pragma unroll
    for (int i = 0; i < W; ++i) {
      const double A = c.NG + H[i];
      d[i] = fma(A, A + c.st, c.ct);
      gam[i] = fma(H[i], fma(c.g2, H[i], c.g1), c.g0);
    }
    div_t<W>(dpv, gam, w);
  } else {
    double den[W], num[W], t[W];
#pragma unroll
    for (int i = 0; i < W; ++i) {
      den[i] = fma(-c.sg, z[i], c.rad_e);
      num[i] = c.rad_e;
    }
    div_t<W>(num, den, t);
#pragma unroll
    for (int i = 0; i < W; ++i) {
      d[i] = fma(t[i], c.inv_r0, c.c0);
      w[i] = dpv[i];
    }
  }
}

// dmax_hi: running maximum of the high words of |d| (compared with the guard once per CTA; NaN and Inf have
// the largest high words, so they trip it too)
template <int MODE, int W>
__device__ __forceinline__ void moment_levels(const double (&z)[W], const double (&dpv)[W], const ColumnC& c,
                                              double (&acc)[NMOMC], unsigned& dmax_hi) {
  double w[W], d[W];
  moment_elements<MODE, W>(z, dpv, c, w, d);
#pragma unroll
  for (int i = 0; i < W; ++i) {
    dmax_hi = max(dmax_hi, (unsigned)__double2hiint(d[i]) & 0x7fffffffu);
    accumulate16(acc, w[i], d[i]);
  }
}

// Per-ring constants of the BC05 moment form, one thread per ring (a few hundred operations that every thread
// of atm_moment_kernel used to repeat for every work unit):
//   [0] a1 c  [1] a2 c  [2] N  [3] c = 1 / (rad_e r0)  [4] lam  [5] mu  [6] g0  [7] g1 / c  [8] g2 / c^2
//   [9] r0  [10] q   (see moment_elements and residual3)
constexpr int RINGC = 12;
__global__ void atm_ring_kernel(int nlat, const double* ring_tab, double rad_e, double hc, double* ringc) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= nlat) return;
  const double inv_rad = 1.0 / rad_e;
  const double* rt = ring_tab + h;
  const double a1 = rt[0], a2 = rt[nlat], N = rt[2 * nlat], N1 = rt[3 * nlat];
  const double sn = rt[4 * nlat], cn = rt[5 * nlat], gs = rt[6 * nlat], c1 = rt[7 * nlat];
  const double ex = N * sn, ez = N1 * cn;
  // expansion point: the ellipsoid's radius at this latitude + hc, in units of rad_e
  const double r0 = sqrt(fma(ez, ez, ex * ex)) * inv_rad + hc * inv_rad;
  // all lengths are scaled by c = 1 / (rad_e r0); the expansion point really used is 1 / (rad_e c):
  // r0 rad_e c = 1 + q,  r0'^n = r0^n (1 + q)^(-n)
  const double cs = 1.0 / (rad_e * r0);
  double qs = residual3(r0, rad_e, cs);
  const double rcs = rad_e * r0;                 // 1 / c
  // e = A (A + lam) + mu takes sin^2 + cos^2 = 1 and a rounded mu: the exact residuals
  //   dk = sin^2 + cos^2 - 1,  dm = (K^2 - 1) - mu   (error-free products and sums)
  // shift e by (dk + dm) for the whole ring, i.e. multiply r^n by 1 + (n / 2)(dk + dm): folded into q
  const double K = -(((N - N1) * cs) * cn);
  const double p1 = sn * sn, e1 = fma(sn, sn, -p1), p2 = cn * cn, e2 = fma(cn, cn, -p2);
  const double sm = p1 + p2, bb = sm - p1;
  const double dk = ((sm - 1.0) + ((p1 - (sm - bb)) + (p2 - bb))) + (e1 + e2);
  const double pk = K * K, ek = fma(K, K, -pk);
  const double mu = pk - 1.0;
  const double dm = (pk - (mu + 1.0)) + ek;
  qs -= 0.5 * (dk + dm);
  double* o = ringc + (size_t)h * RINGC;
  o[0] = a1 * cs;
  o[1] = a2 * cs;
  o[2] = N;
  o[3] = cs;
  o[4] = (2.0 * cn) * K;
  o[5] = mu;
  o[6] = gs;
  o[7] = (-(gs * c1) * inv_rad) * rcs;
  o[8] = ((3.0 * gs) * (inv_rad * inv_rad)) * (rcs * rcs);
  o[9] = r0;
  o[10] = qs;
  o[11] = 0.0;
}

template <int MODE, typename Tin>
__global__ void __launch_bounds__(256, 2)
atm_moment_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmP,
                  const __grid_constant__ AtmMomParams p) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  // A TMA box must start at a 16-byte aligned global address (measured: bench_micro/tma_probe.cu raises
  // "illegal instruction" otherwise), and the mirrored runs start at odd longitudes: the box is BOXA
  // elements wider than the 64 columns and starts at the aligned longitude below the run's first column.
  constexpr int BOXA = 16 / (int)sizeof(Tin);
  constexpr int BOXW = MKCF + BOXA;
  constexpr int BOX_ELEMS = ((MLG * BOXW * (int)sizeof(Tin) + 127) / 128) * 128 / (int)sizeof(Tin);   // 128-byte pitch
  constexpr int STAGE_ELEMS = 2 * 4 * BOX_ELEMS;
  constexpr uint32_t STAGE_BYTES = STAGE_ELEMS * sizeof(Tin);
  constexpr uint32_t TX_BYTES = 2 * 4 * MLG * BOXW * sizeof(Tin);
  Tin* stages = reinterpret_cast<Tin*>(smraw);
  double* Mf = reinterpret_cast<double*>(smraw + MSTAGES * STAGE_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(Mf + 4 * MF_V);
  uint64_t* empty = full + MSTAGES;
  int* box_s = reinterpret_cast<int*>(empty + MSTAGES);      // [nchunks][4] copy of p.box_start

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane >> 3, jj = warp * 8 + (lane & 7);
  const int u_begin = p.cta_unit[blockIdx.x], u_end = p.cta_unit[blockIdx.x + 1];
  if (u_begin >= u_end) return;
  if (tid == 0) {
    for (int s = 0; s < MSTAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < p.nchunks * 4; i += 256) box_s[i] = __ldg(p.box_start + i);
  __syncthreads();

  const int ngroups = p.ngroups, nchunks = p.nchunks;
  const int units_per_field = p.nrings * nchunks;
  const int nitems = (u_end - u_begin) * ngroups;
  // Producer side of the pipeline: item n = (unit, level group) goes to stage n % MSTAGES.  The warp
  // n % 8 issues it (one lane), so the issue cost is spread over the CTA; before refilling a stage it
  // waits until all 8 warps have read the item that was there.
  // The producer's position (field, ring, chunk, level group) of item n is carried incrementally by every
  // thread (uniform, a handful of integer operations per stage) instead of being divided out by the issuing lane.
  int pt, ph, pc, pg = 0;
  {
    const int t = u_begin / units_per_field, rem = u_begin - t * units_per_field;
    pt = t;
    ph = p.h0 + rem / nchunks;
    pc = rem % nchunks;
  }
  auto issue = [&](int n) {
    if (n < nitems && warp == (n & 7) && lane == 0) {
      const int s = n % MSTAGES, k = n / MSTAGES;
      if (k > 0) mbar_wait(&empty[s], (k - 1) & 1);
      const int tf = (p.fields_in_map > 1) ? pt : 0;
      mbar_expect_tx(&full[s], TX_BYTES);
      Tin* dst = stages + (size_t)s * STAGE_ELEMS;
      const int* bs = box_s + pc * 4;
#pragma unroll
      for (int sb = 0; sb < 4; ++sb) {
        const int xa = bs[sb] & ~(BOXA - 1);             // two's complement: floors negative starts too
        tma_load_4d(dst + sb * BOX_ELEMS, &tmZ, &full[s], xa, ph, pg * MLG, tf);
        tma_load_4d(dst + (4 + sb) * BOX_ELEMS, &tmP, &full[s], xa, ph, pg * MLG, tf);
      }
    }
    // advance to item n + 1
    if (++pg == ngroups) {
      pg = 0;
      if (++pc == nchunks) {
        pc = 0;
        if (++ph == p.h0 + p.nrings) {
          ph = p.h0;
          ++pt;
        }
      }
    }
  };
  for (int n = 0; n < MSTAGES - 1; ++n) issue(n);

  int pos = (p.dir[sub] > 0) ? jj : (MKCF - 1 - jj);
  asm volatile("" : "+r"(pos));                      // keep it in a register (not re-derived from the constant bank)
  const int par = warp >> 2, gi = warp & 3;          // contraction role: order parity and group of 16 orders
  double facc[2][2][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) (&facc[0][0][0])[i] = 0.0;
  unsigned dmax_hi = 0;
  int slot = p.slot_base + p.cta_slot[blockIdx.x];
  int item = 0;

  // geoid of this thread's column, fetched one unit ahead
  auto column_of = [&](int u, int& col, double& geo, int& shift) {
    const int t = u / units_per_field, rem = u - t * units_per_field;
    const int h = p.h0 + rem / nchunks, c = rem % nchunks;
    col = __ldg(p.col_grid + c * MKC + tid);
    geo = (col >= 0 && p.geoid != nullptr) ? __ldg(p.geoid + h * p.geo_s0 + col * p.geo_s1) : 0.0;
    shift = box_s[c * 4 + sub] & (BOXA - 1);     // columns between the aligned box start and the run
  };
  int col_next, shift_next;
  double geo_next;
  column_of(u_begin, col_next, geo_next, shift_next);

  for (int u = u_begin; u < u_end; ++u) {
    const int t = u / units_per_field, rem = u - t * units_per_field;
    const int h = p.h0 + rem / nchunks, c = rem % nchunks;
    const bool active = (col_next >= 0);
    const double geo = geo_next;
    // this thread's element of level 0 inside a stage (bytes)
    const unsigned boff = (unsigned)((sub * BOX_ELEMS + shift_next + pos) * (int)sizeof(Tin));
    double sg = 1.0, sp = 1.0;
    if (p.field_scale != nullptr) {
      sg = __ldg(p.field_scale + 2 * (p.t0 + t));
      sp = __ldg(p.field_scale + 2 * (p.t0 + t) + 1);
    }
    ColumnC cc;
    double r0, qs;
    if (MODE == MHB200_METHOD_BC05) {
      const double* rc = p.ringc + (size_t)h * RINGC;
      cc.a1 = __ldg(rc) * sg;
      cc.a2 = __ldg(rc + 1) * (sg * sg);
      cc.cinv = __ldg(rc + 3);
      cc.NG = (__ldg(rc + 2) + geo) * cc.cinv;
      cc.st = __ldg(rc + 4);           // lam
      cc.ct = __ldg(rc + 5);           // mu
      cc.g0 = __ldg(rc + 6);
      cc.g1 = __ldg(rc + 7);
      cc.g2 = __ldg(rc + 8);
      r0 = __ldg(rc + 9);
      qs = __ldg(rc + 10);
      cc.NG1 = 0.0;
      cc.rad_e = cc.sg = cc.inv_r0 = cc.c0 = 0.0;
    } else {
      r0 = 1.0 + p.hc * p.inv_rad;
      cc.rad_e = p.rad_e;
      cc.sg = sg;
      cc.inv_r0 = 1.0 / r0;
      cc.c0 = fma(geo * p.inv_rad, cc.inv_r0, -1.0);
      qs = residual3(r0, 1.0, cc.inv_r0);               // r0 inv_r0 = 1 + q
      cc.a1 = cc.a2 = cc.st = cc.ct = cc.g0 = cc.g1 = cc.g2 = cc.NG = cc.NG1 = cc.cinv = 0.0;
    }

    double acc[NMOMC];
#pragma unroll
    for (int j = 0; j < NMOMC; ++j) acc[j] = 0.0;

    for (int g = 0; g < ngroups; ++g, ++item) {
      issue(item + MSTAGES - 1);
      const int s = item % MSTAGES;
      mbar_wait(&full[s], (item / MSTAGES) & 1);
      const char* sbase = reinterpret_cast<const char*>(stages) + (size_t)s * STAGE_BYTES + boff;
      const Tin* zs = reinterpret_cast<const Tin*>(sbase);
      const Tin* ps = reinterpret_cast<const Tin*>(sbase + 4 * BOX_ELEMS * sizeof(Tin));
      const int nl = p.nlev - g * MLG;
      {
        // (guarded loads + selects on purpose: an unconditional 8-load fast path for full groups measured 6 %
        // slower on the same box -- profiles/README.md, round 2 -- the predicated form lets ptxas spread the
        // loads through the arithmetic); the tail group (nlev % 4 levels) runs narrower instantiations: no
        // padded work
        double z[MLG], dpv[MLG];
#pragma unroll
        for (int k = 0; k < MLG; ++k) {
          z[k] = (k < nl) ? (double)zs[k * BOXW] : 0.0;
          dpv[k] = (k < nl) ? (double)ps[k * BOXW] : 0.0;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (active) {
          if (nl >= MLG) {
            moment_levels<MODE, MLG>(z, dpv, cc, acc, dmax_hi);
          } else if (nl & 2) {
            const double z2[2] = {z[0], z[1]}, d2[2] = {dpv[0], dpv[1]};
            moment_levels<MODE, 2>(z2, d2, cc, acc, dmax_hi);
          }
          if (nl < MLG && (nl & 1)) {
            const double z1[1] = {(nl & 2) ? z[2] : z[0]}, d1[1] = {(nl & 2) ? dpv[2] : dpv[0]};
            moment_levels<MODE, 1>(z1, d1, cc, acc, dmax_hi);
          }
        }
      }
    }

    // first trig fragments of this chunk (L2 latency hides behind the fold) and next unit's column
    const double* Ablk = p.trig + ((size_t)(c * 16) * 16 + (par * 4 + gi) * 2) * 32 + lane;
    double ac[4], as[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ac[i] = __ldg(Ablk + i * 512);
      as[i] = __ldg(Ablk + i * 512 + 32);
    }
    if (u + 1 < u_end) column_of(u + 1, col_next, geo_next, shift_next);

    // fold over the four symmetric columns (phi, -phi, pi - phi, pi + phi), held by the lanes sub = 0..3 of
    // this warp: lane sub = v ends up with fold variant v (0 even-order cosine, 1 even sine, 2 odd cosine,
    // 3 odd sine; see fold_chunk in stokes_grid.cu)
    // (signs as exact +-1 multipliers of FMAs: one operation per stage instead of both branches and a select)
    const double fs1 = (sub & 1) ? -1.0 : 1.0;                       // stage 1: y + fs1 x
    const double fsx = (sub == 2) ? -1.0 : 1.0, fsz = (sub == 1) ? -1.0 : 1.0;     // stage 2: fsx x + fsz zz
#pragma unroll
    for (int j = 0; j < NMOMC; ++j) {
      double x = acc[j];
      const double y = __shfl_xor_sync(0xffffffffu, x, 8);
      x = fma(fs1, x, y);
      const double zz = __shfl_xor_sync(0xffffffffu, x, 16);
      acc[j] = fma(fsz, zz, fsx * x);
    }
    // every warp has finished the previous chunk's contraction (reads of Mf)
    asm volatile("bar.sync 1, 256;" ::: "memory");
    {
      double* dst = Mf + sub * MF_V + (jj >> 2) * MF_KS + (jj & 3);
#pragma unroll
      for (int j = 0; j < NMOMC; ++j) dst[(j >> 3) * 32 + (j & 7) * 4] = acc[j];
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");

    // contraction of the chunk: rows = 8 orders of one parity (cos and sin tiles), k = 4 contraction
    // columns, n = 8 moments; cosine rows read the variant 2*par, sine rows 2*par + 1
    {
      const double* Bc = Mf + (2 * par) * MF_V + lane;
      const double* Bs = Bc + MF_V;
#pragma unroll
      for (int ks = 0; ks < 16; ++ks) {
        const double a_c = ac[ks & 3], a_s = as[ks & 3];
        if (ks + 4 < 16) {
          ac[ks & 3] = __ldg(Ablk + (ks + 4) * 512);
          as[ks & 3] = __ldg(Ablk + (ks + 4) * 512 + 32);
        }
        const double b0 = Bc[ks * MF_KS], b1 = Bc[ks * MF_KS + 32];
        const double b2 = Bs[ks * MF_KS], b3 = Bs[ks * MF_KS + 32];
        dmma884(facc[0][0][0], facc[0][0][1], a_c, b0);
        dmma884(facc[0][1][0], facc[0][1][1], a_c, b1);
        dmma884(facc[1][0][0], facc[1][0][1], a_s, b2);
        dmma884(facc[1][1][0], facc[1][1][1], a_s, b3);
      }
    }

    if (c == nchunks - 1 || u == u_end - 1) {
      // end of this CTA's part of ring h: flush the Fourier sums (sign of einsum(m_phi, -pfactor), :270)
      const double scale = (MODE == MHB200_METHOD_BC05) ? -sp : -sp / 9.80665;
      const int m = 16 * gi + 2 * (lane >> 2) + par;
      double* Fs = p.F + ((size_t)slot * MBLK + m) * 32 + 2 * (lane & 3);
#pragma unroll
      for (int cs = 0; cs < 2; ++cs)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          *reinterpret_cast<double2*>(Fs + cs * 16 + nt * 8) =
              make_double2(facc[cs][nt][0] * scale, facc[cs][nt][1] * scale);
          facc[cs][nt][0] = facc[cs][nt][1] = 0.0;
        }
      if (tid == 0) {
        p.slot_r0[slot] = r0;
        p.slot_q[slot] = qs;
        p.slot_ring[slot] = h;
      }
      ++slot;
    }
  }
  if (__syncthreads_or(dmax_hi > p.dbound_hi) && tid == 0) atomicOr(p.flag, 1);
}

// ------------------------------------------------------------------------------------------------
// K_A, surface pressure (gen_pressure_stokes.py:200-201): any LMAX.
// Unit = (field, group of PNR rings, order block): the moments M_j = (P/G) d^j, d = R / (rad_e r0(h)) - 1,
// of a chunk of 32 contraction columns x PNR rings are folded and contracted against the trig rows of the
// order block; one trig fragment feeds 2 * PNR tiles.  r0(h) is the mid-range of R / rad_e over ring h
// (scanned by the CTA), so |d| is half the ring's radius spread (topography: ~6e-4).
// ------------------------------------------------------------------------------------------------
constexpr int PNR = 4;                    // rings per unit
constexpr int PKCF = 32;                  // contraction columns per chunk
constexpr int PF_V = 8 * MF_KS + 8;       // folded-moment buffer: doubles per fold variant
constexpr int PF_RING = 4 * PF_V;         // ... per ring

struct PressMomParams {
  int nlat, nlon, nchunks, ngroups, nmb, nks_total, Frows;
  const int* col_grid;         // [nchunks][128] grid column of thread column (warp % 4, lane), -1 = inactive
  const double *P, *G, *R;
  long long Ps[3], Gs[3], Rs[3];
  const double* trig;          // A fragments [nmb][nks_total][16 row tiles][32 lanes]
  double* F;                   // [slot = field * nlat + ring][Frows orders][cs][16 moments]
  double* slot_r0;
  double* slot_q;
  int* slot_ring;
  int* flag;
  double rad_e, inv_rad;
  unsigned dbound_hi;
};

// Expansion point of every (field, ring): r0 = mid-range of R / rad_e over the ring, the correction q of the
// point the moment kernel really uses (see residual3), and the slot's ring index.  grid (nlat, nbatch).
__global__ void __launch_bounds__(256) ring_center_kernel(const __grid_constant__ PressMomParams p) {
  __shared__ double red[2][8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int h = blockIdx.x, t = blockIdx.y;
  const double* Rt = p.R + t * p.Rs[0] + h * p.Rs[1];
  double lo = 1.0e300, hi = -1.0e300;
  for (int col = tid; col < p.nlon; col += 256) {
    const double v = __ldg(Rt + col * p.Rs[2]);
    // a NaN radius must not be lost in the comparisons: it poisons both bounds (and then trips the guard)
    lo = (v < lo || v != v) ? v : lo;
    hi = (v > hi || v != v) ? v : hi;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double lo2 = __shfl_xor_sync(0xffffffffu, lo, o), hi2 = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = (lo2 < lo || lo2 != lo2) ? lo2 : lo;
    hi = (hi2 > hi || hi2 != hi2) ? hi2 : hi;
  }
  if (lane == 0) {
    red[0][warp] = lo;
    red[1][warp] = hi;
  }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w) {
      const double lo2 = red[0][w], hi2 = red[1][w];
      lo = (lo2 < lo || lo2 != lo2) ? lo2 : lo;
      hi = (hi2 > hi || hi2 != hi2) ? hi2 : hi;
    }
    const double r0 = (0.5 * (lo + hi)) * p.inv_rad;
    const int slot = t * p.nlat + h;
    p.slot_r0[slot] = r0;
    p.slot_q[slot] = residual3(r0, p.rad_e, 1.0 / (p.rad_e * r0));     // r0 rad_e c = 1 + q
    p.slot_ring[slot] = h;
  }
}

__global__ void __launch_bounds__(256, 2) pressure_moment_kernel(const __grid_constant__ PressMomParams p) {
  extern __shared__ double psm[];
  double* Mf = psm;                              // [PNR][4 variants][PF_V]
  double* ringc = Mf + PNR * PF_RING;            // [PNR] r0, [PNR] 1 / (rad_e r0)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int mb = blockIdx.x % p.nmb, rg = blockIdx.x / p.nmb, t = blockIdx.y;
  const int hbase = rg * PNR;

  // expansion point of every ring of the group (ring_center_kernel): r0 and 1 / (rad_e r0)
  if (tid < PNR) {
    const int h = min(hbase + tid, p.nlat - 1);
    const double r0 = p.slot_r0[t * p.nlat + h];
    ringc[tid] = r0;
    ringc[PNR + tid] = 1.0 / (p.rad_e * r0);
  }
  __syncthreads();

  // produce-phase role: two passes over ring pairs; thread column (warp % 4, lane) of ring 2 pass + tid / 128
  const int sub = lane >> 3, jj = (warp & 3) * 8 + (lane & 7), rhalf = tid >> 7;
  const double fs1 = (sub & 1) ? -1.0 : 1.0, fsx = (sub == 2) ? -1.0 : 1.0, fsz = (sub == 1) ? -1.0 : 1.0;
  // contraction role
  const int par = warp >> 2, gi = warp & 3;
  double facc[PNR][2][2][2];
#pragma unroll
  for (int i = 0; i < PNR * 8; ++i) (&facc[0][0][0][0])[i] = 0.0;
  int viol = 0;

  struct Cell {
    double P, G, R;
    bool ok;
  };
  auto load_cell = [&](int c, int pass) -> Cell {
    Cell x;
    x.P = 0.0;
    x.G = 1.0;
    x.R = 0.0;
    const int h = hbase + 2 * pass + rhalf;
    const int col = (c < p.nchunks) ? __ldg(p.col_grid + c * 128 + (tid & 127)) : -1;
    x.ok = (col >= 0 && h < p.nlat);
    if (x.ok) {
      x.P = __ldg(p.P + t * p.Ps[0] + h * p.Ps[1] + col * p.Ps[2]);
      x.G = __ldg(p.G + t * p.Gs[0] + h * p.Gs[1] + col * p.Gs[2]);
      x.R = __ldg(p.R + t * p.Rs[0] + h * p.Rs[1] + col * p.Rs[2]);
    }
    return x;
  };
  Cell cell[2];
  cell[0] = load_cell(0, 0);
  cell[1] = load_cell(0, 1);

  for (int c = 0; c < p.nchunks; ++c) {
    // first trig fragments of the chunk
    const double* Ablk = p.trig + (((size_t)mb * p.nks_total + c * 8) * 16 + (par * 4 + gi) * 2) * 32 + lane;
    double ac[4], as[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ac[i] = __ldg(Ablk + i * 512);
      as[i] = __ldg(Ablk + i * 512 + 32);
    }
    __syncthreads();          // previous chunk's contraction has read Mf
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int rr = 2 * pass + rhalf;
      double f = 0.0, d = 0.0;
      if (cell[pass].ok) {
        f = cell[pass].P / cell[pass].G;                       // :200, true division
        d = fma(cell[pass].R, ringc[PNR + rr], -1.0);
        viol |= ((unsigned)(__double2hiint(d) & 0x7fffffff) > p.dbound_hi) ? 1 : 0;
      }
      double* dst = Mf + rr * PF_RING + sub * PF_V + (jj >> 2) * MF_KS + (jj & 3);
      double mj = f;
#pragma unroll
      for (int j = 0; j < NMOMC; ++j) {
        // fold over the four symmetric columns (see atm_moment_kernel)
        double x = mj;
        const double y = __shfl_xor_sync(0xffffffffu, x, 8);
        x = fma(fs1, x, y);
        const double zz = __shfl_xor_sync(0xffffffffu, x, 16);
        dst[(j >> 3) * 32 + (j & 7) * 4] = fma(fsz, zz, fsx * x);
        mj *= d;
      }
    }
    __syncthreads();
    // next chunk's cells: their latency hides behind the contraction
    cell[0] = load_cell(c + 1, 0);
    cell[1] = load_cell(c + 1, 1);
    {
      const double* Bc = Mf + (2 * par) * PF_V + lane;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const double a_c = ac[ks & 3], a_s = as[ks & 3];
        if (ks + 4 < 8) {
          ac[ks & 3] = __ldg(Ablk + (ks + 4) * 512);
          as[ks & 3] = __ldg(Ablk + (ks + 4) * 512 + 32);
        }
#pragma unroll
        for (int rr = 0; rr < PNR; ++rr) {
          const double* B = Bc + rr * PF_RING + ks * MF_KS;
          const double b0 = B[0], b1 = B[32], b2 = B[PF_V], b3 = B[PF_V + 32];
          dmma884(facc[rr][0][0][0], facc[rr][0][0][1], a_c, b0);
          dmma884(facc[rr][0][1][0], facc[rr][0][1][1], a_c, b1);
          dmma884(facc[rr][1][0][0], facc[rr][1][0][1], a_s, b2);
          dmma884(facc[rr][1][1][0], facc[rr][1][1][1], a_s, b3);
        }
      }
    }
  }
  // flush: one slot per (field, ring)
  const int m = mb * MBLK + 16 * gi + 2 * (lane >> 2) + par;
#pragma unroll
  for (int rr = 0; rr < PNR; ++rr) {
    const int h = hbase + rr;
    if (h < p.nlat) {
      const int slot = t * p.nlat + h;
      double* Fs = p.F + ((size_t)slot * p.Frows + m) * 32 + 2 * (lane & 3);
#pragma unroll
      for (int cs = 0; cs < 2; ++cs)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
          *reinterpret_cast<double2*>(Fs + cs * 16 + nt * 8) = make_double2(facc[rr][cs][nt][0], facc[rr][cs][nt][1]);
    }
  }
  if (__syncthreads_or(viol) && tid == 0) atomicOr(p.flag, 1);
}

// ------------------------------------------------------------------------------------------------
// K_B: Legendre sum over the slots, per (degree block, order), as DMMA tiles
//   rows = (cs, moment) 32 = 4 row tiles (one per warp), k = 4 slots, n = 8 degrees
// The weighted P_lm values of 128 slots x 16 degrees are generated by the Holmes-Featherstone recursion
// (thread = slot; operation order of gravity_toolkit.plm_holmes as in legendre_ring of stokes_grid.cu)
// into a fragment-major shared-memory tile, double buffered.
// ------------------------------------------------------------------------------------------------
struct LegParams {
  int lmax, mmax, Mpad, Frows, nlb, noff, nsplit, nunits;
  const int* unit_lm;          // [nunits][2]: degree block, order
  const double *f1, *f2, *pmm, *sq, *rescale, *x, *w, *ckpt;
  const double* F;
  const double* slot_r0;
  const double* slot_q;
  const int* slot_ring;
  int field_slot[NB_SUB + 1];
  double* G;
  const int* flag;
};

// Warp-specialised: warps 0-3 run the recursion (thread = slot of the current 128-slot block) and fill a
// double-buffered 16-degree tile; warps 4-7 (row tile = warp - 4) contract it with the block's F fragments.
// The dependent-chain recursion (latency bound) and the DMMA stream (pipe bound) overlap; hand-over by
// named barriers (1, 2: buffer free; 3, 4: buffer full), 128 arrivals + 128 waits each.
__device__ __forceinline__ void named_sync(int id) { asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void named_arrive(int id) { asm volatile("bar.arrive %0, 256;" ::"r"(id) : "memory"); }

__global__ void __launch_bounds__(256, 2) legendre_moment_kernel(const __grid_constant__ LegParams q) {
  if (*reinterpret_cast<const volatile int*>(q.flag) != 0) return;
  __shared__ double tile[2][32 * 64];
  // grid (field x split, unit): CTAs are handed out x-fastest, so every field's longest units (low orders: up to
  // four times the tiles of the high ones) start first and the short ones fill the tail
  const int unit = blockIdx.y, tt = blockIdx.x / q.nsplit, split = blockIdx.x % q.nsplit;
  const int lb = q.unit_lm[2 * unit], m = q.unit_lm[2 * unit + 1];
  const int lbase = lb * LBLK, lend = min(lbase + LBLK - 1, q.lmax);
  const int sb = q.field_slot[tt], se = q.field_slot[tt + 1];
  const int s0 = sb + (int)((long long)(se - sb) * split / q.nsplit);
  const int s1 = sb + (int)((long long)(se - sb) * (split + 1) / q.nsplit);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int first_sub = max(0, m - lbase) >> 4;      // 16-degree sub-blocks wholly below the diagonal are skipped
  const int nsub = ((lend - lbase) >> 4) + 1;
  const int nblocks = (s1 - s0 + 127) / 128;
  const int total = nblocks * (nsub - first_sub);     // tiles handed over

  // Recursion multipliers of this (degree block, order) in shared memory, with seed rows that make the
  // recursion uniform from its first degree: every row is  cur = (x a) p1 - b p2  (operation order of
  // plm_holmes):  l = m: (0, -1) on the state (p2, p1) = (pmm, 0);  l = m + 1: (sqrt(2m + 3), 0);
  // l >= m + 2: (f1, f2).
  __shared__ double2 rec_tab[LBLK];
  if (tid < LBLK) {
    const int l = lbase + tid;
    double2 r = make_double2(0.0, 0.0);
    if (l == m) r = make_double2(0.0, -1.0);
    else if (l == m + 1) r = make_double2(q.sq[m], 0.0);
    else if (l > m + 1 && l <= lend) r = make_double2(q.f1[(size_t)l * q.Mpad + m], q.f2[(size_t)l * q.Mpad + m]);
    rec_tab[tid] = r;
  }
  __syncthreads();

  if (warp < 4) {
    // ---------------- producers: weighted P_lm(x_h) r0(h)^(l + noff) of 128 slots x 16 degrees per tile
    const bool from_ckpt = (q.ckpt != nullptr && lb > 0 && m <= lbase - 2);
    const double pmm = q.pmm[m];
    const int ks = tid >> 2, k = tid & 3;
    const int lstart = lbase + 16 * first_sub;      // first degree of the first tile
    int it = 0;
    for (int blk = 0; blk < nblocks; ++blk) {
      const int s = s0 + blk * 128 + tid;
      const bool valid = (s < s1);
      const int h = valid ? q.slot_ring[s] : 0;
      const double rho = valid ? q.slot_r0[s] : 1.0;
      const double qs = valid ? q.slot_q[s] : 0.0;
      const double xv = q.x[h];
      // (P rs) wh of plm_holmes + weighting, folded into one factor; zero for slots past the end
      const double sc = valid ? q.rescale[(size_t)h * q.Mpad + m] * q.w[h] : 0.0;
      // recursion state (p2, p1) = values at the two degrees below the next one
      double p2, p1;
      if (from_ckpt) {
        const double* ck = q.ckpt + (((size_t)h * q.nlb + lb) * q.Mpad + m) * 2;
        p2 = ck[0];
        p1 = ck[1];
      } else if (m >= lbase) {
        p2 = pmm;                          // the seed row of degree m produces pmm
        p1 = 0.0;
      } else {
        p2 = 0.0;                          // m = lbase - 1: the block starts at degree m + 1
        p1 = pmm;
      }
      // A = sc r0^(l + noff), advanced by r0 per degree and restarted per tile from A0 (advanced by r0^16), so
      // that no value is more than ~20 roundings away from an exact power; nq = (l + noff) q is the
      // correction of the expansion point K_A really used (see residual3)
      const double rho2 = rho * rho, rho4 = rho2 * rho2, rho8 = rho4 * rho4, rho16 = rho8 * rho8;
      double A0 = sc * ipow(rho, lstart + q.noff);
      for (int sub = first_sub; sub < nsub; ++sub, ++it) {
        const int b = it & 1;
        double* T = tile[b] + ks * 64 + k;
        const int l0 = lbase + 16 * sub;
        double A = A0;
        double nq = (double)(l0 + q.noff) * qs;
        named_sync(1 + b);                 // consumers have released this buffer
#pragma unroll
        for (int dl = 0; dl < 16; ++dl) {
          const int l = l0 + dl;
          double v = 0.0;
          if (l >= m && l <= lend) {       // uniform over the CTA
            const double2 ab = rec_tab[l - lbase];
            const double cur = __dsub_rn(__dmul_rn(__dmul_rn(xv, ab.x), p1), __dmul_rn(ab.y, p2));
            p2 = p1;
            p1 = cur;
            v = cur * fma(-nq, A, A);
          }
          A *= rho;
          nq += qs;
          T[(dl >> 3) * 32 + ((((dl & 7) + ks) & 7) << 2)] = v;
        }
        named_arrive(3 + b);
        A0 *= rho16;
      }
    }
  } else {
    // ---------------- consumers: acc[(cs, j) rows of this warp][degrees] += F^T (rows x slots) * tile
    const int cw = warp - 4;
    double acc[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = 0.0;
    if (total >= 1) named_arrive(1);
    if (total >= 2) named_arrive(2);
    const int n = lane >> 2;
    int it = 0;
    for (int blk = 0; blk < nblocks; ++blk) {
      const int sblk = s0 + blk * 128;
      // A fragments of the block: F[slot][m][(cs, j)], row tile = cw
      double af[32];
#pragma unroll
      for (int kk = 0; kk < 32; ++kk) {
        const int s = sblk + 4 * kk + (lane & 3);
        af[kk] = (s < s1) ? __ldg(q.F + ((size_t)s * q.Frows + m) * 32 + cw * 8 + (lane >> 2)) : 0.0;
      }
      // pull the next block's rows (256 bytes per slot) into L2 while this one is contracted
      {
        const int s = sblk + 128 + (tid - 128);
        if (s < s1) {
          const double* nx = q.F + ((size_t)s * q.Frows + m) * 32;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 16));
        }
      }
#pragma unroll
      for (int sub = 0; sub < 4; ++sub) {
        if (sub >= first_sub && sub < nsub) {
          const int b = it & 1;
          named_sync(3 + b);               // tile is full
          const double* Tb = tile[b] + (lane & 3);
#pragma unroll
          for (int kk = 0; kk < 32; ++kk) {
            const double b0 = Tb[kk * 64 + (((n + kk) & 7) << 2)];
            const double b1 = Tb[kk * 64 + 32 + (((n + kk) & 7) << 2)];
            dmma884(acc[2 * sub][0], acc[2 * sub][1], af[kk], b0);
            dmma884(acc[2 * sub + 1][0], acc[2 * sub + 1][1], af[kk], b1);
          }
          if (it + 2 < total) named_arrive(1 + b);     // the last two tiles have no successor
          ++it;
        }
      }
    }
    double* Gd = q.G + (((size_t)(tt * q.nsplit + split) * q.nunits + unit) * 4 + cw) * 512 + lane * 2;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) *reinterpret_cast<double2*>(Gd + nt * 64) = make_double2(acc[nt][0], acc[nt][1]);
  }
}

// ------------------------------------------------------------------------------------------------
// K_C: binomial expansion + degree factors:  C_lm = dfactor[l] sum_j C(l + n0, j) sum_split G
// ------------------------------------------------------------------------------------------------
struct FinParams {
  int lmax, mmax, nsplit, nunits;
  const int* lb_unit0;         // first unit of every degree block
  const double* E;             // [16][lmax + 1] binomial coefficients C(l + n0, j)
  const double* G;
  double *clm, *slm;
  const int* flag;
  const double* dfactor_dev;   // used when lmax + 1 > 448
  double dfactor[448];
};

__global__ void __launch_bounds__(64) moment_finalize_kernel(const __grid_constant__ FinParams q) {
  if (*reinterpret_cast<const volatile int*>(q.flag) != 0) return;
  const int l = blockIdx.x, tt = blockIdx.z;
  const double df = (q.dfactor_dev != nullptr) ? q.dfactor_dev[l] : q.dfactor[l];
  const int nt = (l & 63) >> 3, ll = (l & 7) >> 1, e = l & 1;
  for (int m = blockIdx.y * blockDim.x + threadIdx.x; m <= q.mmax; m += gridDim.y * blockDim.x) {
    double c = 0.0, s = 0.0;
    if (m <= l) {
      const int unit = q.lb_unit0[l >> 6] + m;
      for (int sp = 0; sp < q.nsplit; ++sp) {
        const double* g = q.G + ((size_t)(tt * q.nsplit + sp) * q.nunits + unit) * G_UNIT + nt * 64 + ll * 2 + e;
        double cc = 0.0, ss = 0.0;
#pragma unroll
        for (int j = NMOMC - 1; j >= 0; --j) {          // small terms first
          const double ej = __ldg(q.E + (size_t)j * (q.lmax + 1) + l);
          const int o = (j >> 3) * 512 + (j & 7) * 8;   // row tile j / 8 (+2 for sine), lane (j % 8) * 4 + ll
          cc = fma(ej, g[o], cc);
          ss = fma(ej, g[o + 1024], ss);
        }
        c += cc;
        s += ss;
      }
    }
    const size_t o = ((size_t)tt * (q.lmax + 1) + l) * (q.mmax + 1) + m;
    q.clm[o] = c * df;
    q.slm[o] = s * df;
  }
}

// ================================================================================================
// host side
// ================================================================================================
struct MomSchedule {
  int nb = 0, h0 = 0, h1 = 0, ncta = 0, nslots = 0;
  int* cta_unit_dev = nullptr;
  int* cta_slot_dev = nullptr;
  std::vector<int> field_slot;          // [nb + 1]
};

struct MomentPlan {
  bool tma_ok = false;                  // the fold sub-blocks of every chunk are contiguous longitude runs
  int nchunks = 0, nf = 0, nunits = 0, ncta_max = 0;
  int dir[4] = {1, 1, 1, 1};
  int *box_start = nullptr, *col_grid = nullptr, *unit_lm = nullptr, *lb_unit0 = nullptr;
  int nchunks32 = 0;
  int* col_grid32 = nullptr;            // pressure kernel: [nchunks32][128]
  double* trig = nullptr;               // [nmb][nchunks * 16][16][32]
  double* E = nullptr;                  // [2][16][lmax + 1]: n0 = 2, n0 = 4
  std::vector<MomSchedule> sched;
};

namespace {

template <typename T>
int upload_vec(T** dev, const std::vector<T>& host) {
  MHB_CUDA(cudaMalloc((void**)dev, std::max<size_t>(host.size(), 1) * sizeof(T)));
  if (!host.empty()) MHB_CUDA(cudaMemcpy(*dev, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
  return MHB200_OK;
}

size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }

// layout of the moment region of a workspace
struct MomWs {
  size_t flag, ringc, slot_r0, slot_q, slot_ring, F, G, total;
};
MomWs mom_layout(const mhb200_plan* pl, size_t max_slots, size_t g_units) {
  MomWs w;
  size_t o = 0;
  w.flag = o;
  o += 256;
  w.ringc = o;
  o = align256(o + (size_t)pl->nlat * RINGC * sizeof(double));
  w.slot_r0 = o;
  o = align256(o + max_slots * sizeof(double));
  w.slot_q = o;
  o = align256(o + max_slots * sizeof(double));
  w.slot_ring = o;
  o = align256(o + max_slots * sizeof(int));
  w.F = o;
  o = align256(o + max_slots * (size_t)pl->Mpad * 32 * sizeof(double));
  w.G = o;
  o = align256(o + g_units * G_UNIT * sizeof(double));
  w.total = o;
  return w;
}

// fields per K_A launch: at most NB_SUB, fewer when a slot is large (LMAX 360: 98 KB per ring)
int nb_sub(const mhb200_plan* pl) {
  const double per_field = (double)pl->nlat * pl->Mpad * 32.0 * sizeof(double);
  return std::max(1, std::min(NB_SUB, (int)(256.0e6 / per_field)));
}

int pick_nsplit(const mhb200_plan* pl, int nb) {
  const int want = (2 * pl->num_sms + nb * pl->mom->nunits - 1) / (nb * pl->mom->nunits);
  return std::max(1, std::min(MAX_SPLIT, want));
}

size_t g_units_bound(const mhb200_plan* pl, int nb_max) {
  size_t best = 0;
  for (int nb = 1; nb <= nb_max; ++nb)
    best = std::max(best, (size_t)nb * pick_nsplit(pl, nb) * pl->mom->nunits);
  return best;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) != cudaSuccess ||
        qr != cudaDriverEntryPointSuccess)
      f = nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}

// 4-D map (lon, lat, level, field) of a (field, level, lat, lon) cube; box = 64 lon x 1 lat x MLG levels
int make_cube_map(CUtensorMap* tm, const void* base, int dtype, int nlon, int nlat, int nlev, int nfields,
                  long long batch_stride) {
  EncodeTiledFn enc = tensor_map_encoder();
  if (!enc) {
    set_err("cuTensorMapEncodeTiled is not available from the driver");
    return MHB200_ECUDA;
  }
  const size_t esz = (dtype == MHB200_F64) ? 8 : 4;
  cuuint64_t dims[4] = {(cuuint64_t)nlon, (cuuint64_t)nlat, (cuuint64_t)nlev, (cuuint64_t)std::max(nfields, 1)};
  cuuint64_t strides[3] = {(cuuint64_t)nlon * esz, (cuuint64_t)nlon * nlat * esz,
                           (cuuint64_t)(nfields > 1 ? batch_stride : (long long)nlon * nlat * nlev) * esz};
  cuuint32_t box[4] = {(cuuint32_t)(MKCF + 16 / esz), 1, (cuuint32_t)MLG, 1};   // see atm_moment_kernel (BOXW)
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(tm, dtype == MHB200_F64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                         const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_err("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return MHB200_ECUDA;
  }
  return MHB200_OK;
}

MomSchedule* get_schedule(mhb200_plan* pl, int nb, int h0, int h1, int* rc_out) {
  MomentPlan* mp = pl->mom;
  *rc_out = MHB200_OK;
  for (auto& s : mp->sched)
    if (s.nb == nb && s.h0 == h0 && s.h1 == h1) return &s;
  MomSchedule sc;
  sc.nb = nb;
  sc.h0 = h0;
  sc.h1 = h1;
  std::vector<int> cta_unit, cta_slot, field_slot;
  sc.ncta = moment_schedule(nb, h1 - h0, mp->nchunks, mp->ncta_max, cta_unit, cta_slot, field_slot, &sc.nslots);
  sc.field_slot = field_slot;
  int rc;
  if ((rc = upload_vec(&sc.cta_unit_dev, cta_unit)) != MHB200_OK || (rc = upload_vec(&sc.cta_slot_dev, cta_slot)) != MHB200_OK) {
    *rc_out = rc;
    return nullptr;
  }
  mp->sched.push_back(sc);
  return &mp->sched.back();
}

}  // namespace

// Equal split of the linearised (field, ring, chunk) units over the CTAs; a slot is the part of one ring a
// CTA owns.  Pure host logic (exposed to the CPU tests through mhb200_moment_schedule_probe).
int moment_schedule(int nb, int nrings, int nchunks, int max_ctas, std::vector<int>& cta_unit,
                    std::vector<int>& cta_slot, std::vector<int>& field_slot, int* nslots) {
  const long long units = (long long)nb * nrings * nchunks;
  const int ncta = (int)std::min<long long>(max_ctas, units);
  cta_unit.assign(ncta + 1, 0);
  cta_slot.assign(ncta, 0);
  field_slot.assign(nb + 1, 0);
  for (int b = 0; b <= ncta; ++b) cta_unit[b] = (int)(units * b / ncta);
  int slot = 0, last_field = -1;
  for (int b = 0; b < ncta; ++b) {
    cta_slot[b] = slot;
    long long prev_ring = -1;
    for (int u = cta_unit[b]; u < cta_unit[b + 1]; ++u) {
      const long long ring = u / nchunks;          // (field, ring) linearised
      if (ring != prev_ring) {
        const int t = (int)(ring / nrings);
        while (last_field < t) field_slot[++last_field] = slot;
        ++slot;
        prev_ring = ring;
      }
    }
  }
  while (last_field < nb) field_slot[++last_field] = slot;
  *nslots = slot;
  return ncta;
}

// Pure host logic (CPU-tested through mhb200_moment_layout_probe): thread-column map and TMA box starts of
// the atmosphere kernel's chunks for the four fold sub-blocks sets[0..3] (grid columns at phi, -phi, pi - phi,
// pi + phi of every contraction column).
//   thread column (warp, lane) of chunk c: sub = lane / 8, contraction column j = 64 c + 8 warp + lane % 8.
//   A grid column that appears in two sub-blocks of the same contraction column (phi = 0, pi, +-pi/2) is kept
//   in the first and dropped from the others, so every trig entry has weight 1.
//   Returns whether every sub-block of every chunk is a contiguous longitude run (ascending or descending with
//   the contraction column, the same direction in all chunks) -- the condition for the TMA feed.
bool moment_layout(int nf, const std::vector<int>* const sets[4], int dir[4], std::vector<int>& col_grid,
                   std::vector<int>& box_start) {
  const int nchunks = (nf + MKCF - 1) / MKCF;
  col_grid.assign((size_t)nchunks * MKC, -1);
  box_start.assign((size_t)nchunks * 4, 0);
  auto column = [&](int sb, int j) -> int {
    if (j >= nf) return -1;
    const int col = (*sets[sb])[j];
    for (int s2 = 0; s2 < sb; ++s2)
      if ((*sets[s2])[j] == col) return -1;
    return col;
  };
  // direction of every sub-block's run (from the first two contraction columns that are both kept)
  bool ok = true;
  for (int sb = 0; sb < 4; ++sb) {
    int d = 0;
    for (int j = 0; j + 1 < nf && d == 0; ++j) {
      const int a = column(sb, j), b = column(sb, j + 1);
      if (a >= 0 && b >= 0) d = b - a;
    }
    if (d == 0) d = 1;                       // a single column
    if (d != 1 && d != -1) ok = false;
    dir[sb] = (d < 0) ? -1 : 1;
  }
  for (int c = 0; c < nchunks; ++c)
    for (int sb = 0; sb < 4; ++sb) {
      bool have = false;
      int start = 0;
      for (int jj = 0; jj < MKCF; ++jj) {
        const int col = column(sb, c * MKCF + jj);
        const int warp = jj >> 3, lane = sb * 8 + (jj & 7);
        col_grid[(size_t)c * MKC + warp * 32 + lane] = col;
        if (col < 0) continue;
        const int pos = (dir[sb] > 0) ? jj : (MKCF - 1 - jj);
        if (!have) {
          start = col - pos;
          have = true;
        } else if (col - pos != start) {
          ok = false;
        }
      }
      box_start[(size_t)c * 4 + sb] = have ? start : 0;
    }
  return ok;
}

int moment_plan_create(mhb200_plan* pl, const double* phi) {
  pl->mom = nullptr;
  if (pl->fold != 2) return MHB200_OK;
  MomentPlan* mp = new MomentPlan();
  pl->mom = mp;
  const int nf = (int)pl->colp.size();
  mp->nf = nf;
  mp->nchunks = (nf + MKCF - 1) / MKCF;
  mp->ncta_max = 2 * pl->num_sms;
  const std::vector<int>* sets[4] = {&pl->colp, &pl->colq, &pl->colp2, &pl->colq2};
  std::vector<int> col_grid, box_start;
  const bool ok = moment_layout(nf, sets, mp->dir, col_grid, box_start);
  auto column = [&](int sb, int j) -> int {
    if (j >= nf) return -1;
    const int col = (*sets[sb])[j];
    for (int s2 = 0; s2 < sb; ++s2)
      if ((*sets[s2])[j] == col) return -1;
    return col;
  };
  mp->tma_ok = ok;
  // pressure kernel: chunks of 32 contraction columns, thread column (warp % 4, lane)
  mp->nchunks32 = (nf + PKCF - 1) / PKCF;
  std::vector<int> col_grid32((size_t)mp->nchunks32 * 128, -1);
  for (int c = 0; c < mp->nchunks32; ++c)
    for (int sb = 0; sb < 4; ++sb)
      for (int jj = 0; jj < PKCF; ++jj)
        col_grid32[(size_t)c * 128 + (jj >> 3) * 32 + sb * 8 + (jj & 7)] = column(sb, c * PKCF + jj);
  // trig table (m_phi of gen_pressure_stokes.py:176-177 on the contraction columns), fragment-major:
  // row tile (par*4 + gi)*2 + {cos, sin}, row r: order 64 mb + 16 gi + 2 r + par; k: column 4 ks + lane % 4
  const int nks = mp->nchunks * 16;
  std::vector<double> trig((size_t)pl->nmb * nks * 512, 0.0);
  for (int mb = 0; mb < pl->nmb; ++mb)
    for (int ks = 0; ks < nks; ++ks)
      for (int g = 0; g < 8; ++g)
        for (int lane = 0; lane < 32; ++lane) {
          const int par = g >> 2, gi = g & 3;
          const int m = mb * MBLK + 16 * gi + 2 * (lane >> 2) + par;
          const int j = ks * 4 + (lane & 3);
          if (m > pl->mmax || j >= nf) continue;
          const double arg = (double)m * phi[pl->colp[j]];
          const size_t base = (((size_t)mb * nks + ks) * 16 + g * 2) * 32 + lane;
          trig[base] = cos(arg);
          trig[base + 32] = sin(arg);
        }
  // K_B units: (degree block, order), orders 0 .. min(64 lb + 63, mmax)
  std::vector<int> unit_lm, lb_unit0(pl->nlb, 0);
  for (int lb = 0; lb < pl->nlb; ++lb) {
    lb_unit0[lb] = (int)unit_lm.size() / 2;
    for (int m = 0; m <= std::min(lb * LBLK + LBLK - 1, pl->mmax); ++m) {
      unit_lm.push_back(lb);
      unit_lm.push_back(m);
    }
  }
  mp->nunits = (int)unit_lm.size() / 2;
  // expansion coefficients C(a, j) = a (a - 1) ... (a - j + 1) / j!:
  //   table 0: a = l + 2 (pressure, expansion in d)    table 1: a = l + 4 (SW02, in d)
  //   table 2: a = (l + 2) / 2, generalised (BC05, expansion in e = (1 + d)^2 - 1)
  const int nl = pl->lmax + 1;
  std::vector<double> E((size_t)3 * NMOMC * nl, 0.0);
  for (int v = 0; v < 3; ++v)
    for (int l = 0; l < nl; ++l) {
      const long double a = (v == 0) ? (long double)(l + 2) : (v == 1) ? (long double)(l + 4) : 0.5L * (l + 2);
      long double caj = 1.0L;
      for (int j = 0; j < NMOMC; ++j) {
        if (j > 0) caj = caj * (a - (long double)(j - 1)) / (long double)j;
        E[((size_t)v * NMOMC + j) * nl + l] = (double)caj;
      }
    }
  int rc;
  if ((rc = upload_vec(&mp->col_grid, col_grid)) || (rc = upload_vec(&mp->box_start, box_start)) ||
      (rc = upload_vec(&mp->col_grid32, col_grid32)) ||
      (rc = upload_vec(&mp->trig, trig)) || (rc = upload_vec(&mp->unit_lm, unit_lm)) ||
      (rc = upload_vec(&mp->lb_unit0, lb_unit0)) || (rc = upload_vec(&mp->E, E)))
    return rc;
  return MHB200_OK;
}

void moment_plan_destroy(mhb200_plan* pl) {
  MomentPlan* mp = pl->mom;
  if (!mp) return;
  cudaFree(mp->box_start);
  cudaFree(mp->col_grid);
  cudaFree(mp->col_grid32);
  cudaFree(mp->unit_lm);
  cudaFree(mp->lb_unit0);
  cudaFree(mp->trig);
  cudaFree(mp->E);
  for (auto& s : mp->sched) {
    cudaFree(s.cta_unit_dev);
    cudaFree(s.cta_slot_dev);
  }
  delete mp;
  pl->mom = nullptr;
}

// bytes of the moment region for calls with up to nb_max fields per launch and `launches` K_A launches
size_t moment_workspace_bytes(const mhb200_plan* pl, int nbatch, int launches) {
  if (!pl->mom) return 0;
  const int nb = std::min(nbatch, nb_sub(pl));
  const size_t max_slots = (size_t)nb * pl->nlat + (size_t)launches * pl->mom->ncta_max;
  return mom_layout(pl, max_slots, g_units_bound(pl, nb)).total;
}

static int moment_mode_env() {
  // MHB200_MOMENTS: 0 direct powers, 1 moments expanded before the Fourier sum (stokes_grid.cu),
  // 2 (default) contracted moments (this file)
  if (const char* e = getenv("MHB200_MOMENTS")) return atoi(e);
  return 2;
}

bool moment_atm_eligible(const mhb200_plan* pl, int dtype, const void* GPH, const void* dP, long long batch_stride,
                         int nbatch, const double* plm_table) {
  const MomentPlan* mp = pl->mom;
  if (!mp || !mp->tma_ok || pl->nlb != 1 || plm_table != nullptr || moment_mode_env() < 2) return false;
  if (!tensor_map_encoder()) return false;
  const size_t esz = (dtype == MHB200_F64) ? 8 : 4;
  if (((uintptr_t)GPH | (uintptr_t)dP) & 15) return false;
  if (((size_t)pl->nlon * esz) % 16) return false;
  if (nbatch > 1 && batch_stride != 0 && ((size_t)batch_stride * esz) % 16) return false;
  return true;
}

struct MomCall {
  int mode, dtype, nlev;
  const void *GPH, *dP;
  long long batch_stride;
  const double* field_scale_dev;
  const double* geoid;
  long long geo_s0, geo_s1;
  const double* ring_tab;
  double rad_e;
};

static int launch_atm_kernel(const MomCall& c, const CUtensorMap& tz, const CUtensorMap& tp, const AtmMomParams& ap,
                             int ncta, cudaStream_t st) {
  const size_t esz = (c.dtype == MHB200_F64) ? 8 : 4;
  const size_t box_bytes = ((MLG * (MKCF + 16 / esz) * esz + 127) / 128) * 128;
  const size_t smem = (size_t)MSTAGES * 2 * 4 * box_bytes + 4 * MF_V * sizeof(double) + 2 * MSTAGES * 8 +
                      (size_t)ap.nchunks * 4 * sizeof(int);
#define MHB_LAUNCH_ATM(MODE, T)                                                                              \
  do {                                                                                                       \
    auto kern = atm_moment_kernel<MODE, T>;                                                                  \
    MHB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
    profile_begin(st);                                                                                       \
    kern<<<ncta, 256, smem, st>>>(tz, tp, ap);                                                               \
    profile_end(st);                                                                                         \
  } while (0)
  if (c.mode == MHB200_METHOD_BC05) {
    if (c.dtype == MHB200_F64) MHB_LAUNCH_ATM(MHB200_METHOD_BC05, double);
    else MHB_LAUNCH_ATM(MHB200_METHOD_BC05, float);
  } else {
    if (c.dtype == MHB200_F64) MHB_LAUNCH_ATM(MHB200_METHOD_SW02, double);
    else MHB_LAUNCH_ATM(MHB200_METHOD_SW02, float);
  }
#undef MHB_LAUNCH_ATM
  g_launches.fetch_add(1);
  MHB_CUDA(cudaGetLastError());
  return MHB200_OK;
}

// K_A for fields [t0, t0 + nb) and rings [h0, h1); slots start at slot_base.  *nslots_out = slots written.
static int moment_atm_launch(mhb200_plan* pl, const MomCall& c, const MomWs& lay, char* ws, int t0, int nb, int h0,
                             int h1, int slot_base, MomSchedule** sched_out, cudaStream_t st) {
  MomentPlan* mp = pl->mom;
  int rc;
  MomSchedule* sc = get_schedule(pl, nb, h0, h1, &rc);
  if (!sc) return rc;
  *sched_out = sc;
  const size_t esz = (c.dtype == MHB200_F64) ? 8 : 4;
  const bool shared = (c.batch_stride == 0);
  const char* gbase = (const char*)c.GPH + (shared ? 0 : (size_t)t0 * c.batch_stride * esz);
  const char* pbase = (const char*)c.dP + (shared ? 0 : (size_t)t0 * c.batch_stride * esz);
  CUtensorMap tz, tp;
  const int nf_map = shared ? 1 : nb;
  if ((rc = make_cube_map(&tz, gbase, c.dtype, pl->nlon, pl->nlat, c.nlev, nf_map, c.batch_stride)) != MHB200_OK) return rc;
  if ((rc = make_cube_map(&tp, pbase, c.dtype, pl->nlon, pl->nlat, c.nlev, nf_map, c.batch_stride)) != MHB200_OK) return rc;
  AtmMomParams ap;
  memset(&ap, 0, sizeof(ap));
  ap.nlat = pl->nlat;
  ap.nlon = pl->nlon;
  ap.nlev = c.nlev;
  ap.nchunks = mp->nchunks;
  ap.ngroups = (c.nlev + MLG - 1) / MLG;
  ap.h0 = h0;
  ap.nrings = h1 - h0;
  ap.t0 = t0;
  ap.fields_in_map = nf_map;
  ap.cta_unit = sc->cta_unit_dev;
  ap.cta_slot = sc->cta_slot_dev;
  ap.slot_base = slot_base;
  ap.box_start = mp->box_start;
  ap.col_grid = mp->col_grid;
  for (int i = 0; i < 4; ++i) ap.dir[i] = mp->dir[i];
  ap.ring_tab = c.ring_tab;
  ap.ringc = (const double*)(ws + lay.ringc);
  ap.geoid = c.geoid;
  ap.geo_s0 = c.geo_s0;
  ap.geo_s1 = c.geo_s1;
  ap.field_scale = c.field_scale_dev;
  ap.trig = mp->trig;
  ap.F = (double*)(ws + lay.F);
  ap.slot_r0 = (double*)(ws + lay.slot_r0);
  ap.slot_q = (double*)(ws + lay.slot_q);
  ap.slot_ring = (int*)(ws + lay.slot_ring);
  ap.flag = (int*)(ws + lay.flag);
  ap.rad_e = c.rad_e;
  ap.inv_rad = 1.0 / c.rad_e;
  ap.hc = MOM_CENTER_HEIGHT;
  // BC05: (n_max / 2) |e| <= 0.8, n = l + 2;  SW02: n_max |d| <= 0.8, n = l + 4
  const double bound = (c.mode == MHB200_METHOD_BC05) ? 2.0 * MOM_ND_BOUND / (double)(pl->lmax + 2)
                                                      : MOM_ND_BOUND / (double)(pl->lmax + 4);
  unsigned long long bits;
  memcpy(&bits, &bound, 8);
  ap.dbound_hi = (unsigned)(bits >> 32);
  return launch_atm_kernel(c, tz, tp, ap, sc->ncta, st);
}

// K_B + K_C over the slots described by field_slot[0..nb]
// noff: r0(h)^(l + noff) in K_B; etable: expansion coefficients of K_C (see moment_plan_create)
static int moment_reduce(mhb200_plan* pl, const MomWs& lay, char* ws, int nb, const int* field_slot, int noff,
                         int etable,
                         const double* dfactor, const double* dfactor_dev, double* clm, double* slm,
                         cudaStream_t st) {
  MomentPlan* mp = pl->mom;
  const int nsplit = pick_nsplit(pl, nb);
  LegParams q;
  memset(&q, 0, sizeof(q));
  q.lmax = pl->lmax;
  q.mmax = pl->mmax;
  q.Mpad = pl->Mpad;
  q.Frows = pl->Mpad;
  q.nlb = pl->nlb;
  q.noff = noff;
  q.nsplit = nsplit;
  q.nunits = mp->nunits;
  q.unit_lm = mp->unit_lm;
  q.f1 = pl->f1;
  q.f2 = pl->f2;
  q.pmm = pl->pmm;
  q.sq = pl->sq;
  q.rescale = pl->rescale;
  q.x = pl->x;
  q.w = pl->w;
  q.ckpt = pl->ckpt;
  q.F = (const double*)(ws + lay.F);
  q.slot_r0 = (const double*)(ws + lay.slot_r0);
  q.slot_q = (const double*)(ws + lay.slot_q);
  q.slot_ring = (const int*)(ws + lay.slot_ring);
  for (int i = 0; i <= nb; ++i) q.field_slot[i] = field_slot[i];
  q.G = (double*)(ws + lay.G);
  q.flag = (const int*)(ws + lay.flag);
  legendre_moment_kernel<<<dim3(nb * nsplit, mp->nunits), 256, 0, st>>>(q);
  g_launches.fetch_add(1);
  MHB_CUDA(cudaGetLastError());
  FinParams f;
  f.lmax = pl->lmax;
  f.mmax = pl->mmax;
  f.nsplit = nsplit;
  f.nunits = mp->nunits;
  f.lb_unit0 = mp->lb_unit0;
  f.E = mp->E + (size_t)etable * NMOMC * (pl->lmax + 1);
  f.G = q.G;
  f.clm = clm;
  f.slm = slm;
  f.flag = q.flag;
  f.dfactor_dev = dfactor_dev;
  if (!dfactor_dev)
    for (int l = 0; l <= pl->lmax; ++l) f.dfactor[l] = dfactor[l];
  moment_finalize_kernel<<<dim3(pl->lmax + 1, (pl->mmax + 64) / 64, nb), 64, 0, st>>>(f);
  g_launches.fetch_add(1);
  MHB_CUDA(cudaGetLastError());
  return MHB200_OK;
}

// Device-resident cubes, all rings: sub-batches of NB_SUB fields through K_A, K_B, K_C.
// `mom_ws` is the moment region of the caller's workspace; *flag_out = device flag the fallback reads.
int moment_atm_run(mhb200_plan* pl, int mode, int dtype, const void* GPH, const void* dP, long long batch_stride,
                   const double* field_scale_dev, const double* geoid, long long geo_s0, long long geo_s1,
                   const double* ring_tab, int nlev, int nbatch, double rad_e, const double* dfactor,
                   double* clm, double* slm, void* mom_ws, const int** flag_out, cudaStream_t st) {
  const int nbmax = std::min(nbatch, nb_sub(pl));
  const MomWs lay = mom_layout(pl, (size_t)nbmax * pl->nlat + pl->mom->ncta_max, g_units_bound(pl, nbmax));
  char* ws = (char*)mom_ws;
  MHB_CUDA(cudaMemsetAsync(ws + lay.flag, 0, sizeof(int), st));
  *flag_out = (const int*)(ws + lay.flag);
  if (mode == MHB200_METHOD_BC05) {
    atm_ring_kernel<<<(pl->nlat + 127) / 128, 128, 0, st>>>(pl->nlat, ring_tab, rad_e, MOM_CENTER_HEIGHT,
                                                            (double*)(ws + lay.ringc));
    g_launches.fetch_add(1);
    MHB_CUDA(cudaGetLastError());
  }
  MomCall c;
  c.mode = mode;
  c.dtype = dtype;
  c.nlev = nlev;
  c.GPH = GPH;
  c.dP = dP;
  c.batch_stride = batch_stride;
  c.field_scale_dev = field_scale_dev;
  c.geoid = geoid;
  c.geo_s0 = geo_s0;
  c.geo_s1 = geo_s1;
  c.ring_tab = ring_tab;
  c.rad_e = rad_e;
  const size_t nout = (size_t)(pl->lmax + 1) * (pl->mmax + 1);
  const int noff = (mode == MHB200_METHOD_BC05) ? 2 : 4;
  for (int t0 = 0; t0 < nbatch; t0 += nbmax) {
    const int nb = std::min(nbmax, nbatch - t0);
    MomSchedule* sc = nullptr;
    int rc = moment_atm_launch(pl, c, lay, ws, t0, nb, 0, pl->nlat, 0, &sc, st);
    if (rc != MHB200_OK) return rc;
    rc = moment_reduce(pl, lay, ws, nb, sc->field_slot.data(), noff, (noff == 2) ? 2 : 1, dfactor, nullptr, clm + t0 * nout,
                       slm + t0 * nout, st);
    if (rc != MHB200_OK) return rc;
  }
  return MHB200_OK;
}

// Host-band path: one field, K_A per latitude band (as its H2D copy lands), one K_B/K_C at the end.
int moment_band_begin(mhb200_plan* pl, int nbands, int mode, const double* ring_tab, double rad_e, void* mom_ws,
                      const int** flag_out, cudaStream_t st) {
  const MomWs lay = mom_layout(pl, (size_t)pl->nlat + (size_t)nbands * pl->mom->ncta_max, g_units_bound(pl, 1));
  MHB_CUDA(cudaMemsetAsync((char*)mom_ws + lay.flag, 0, sizeof(int), st));
  *flag_out = (const int*)((char*)mom_ws + lay.flag);
  if (mode == MHB200_METHOD_BC05) {
    atm_ring_kernel<<<(pl->nlat + 127) / 128, 128, 0, st>>>(pl->nlat, ring_tab, rad_e, MOM_CENTER_HEIGHT,
                                                            (double*)((char*)mom_ws + lay.ringc));
    g_launches.fetch_add(1);
    MHB_CUDA(cudaGetLastError());
  }
  return MHB200_OK;
}

int moment_band_launch(mhb200_plan* pl, int nbands, int mode, int dtype, const void* GPH, const void* dP,
                       const double* geoid, long long geo_s0, long long geo_s1, const double* ring_tab, int nlev,
                       double rad_e, int h0, int h1, int* slot_base, void* mom_ws, cudaStream_t st) {
  const MomWs lay = mom_layout(pl, (size_t)pl->nlat + (size_t)nbands * pl->mom->ncta_max, g_units_bound(pl, 1));
  MomCall c;
  c.mode = mode;
  c.dtype = dtype;
  c.nlev = nlev;
  c.GPH = GPH;
  c.dP = dP;
  c.batch_stride = 0;
  c.field_scale_dev = nullptr;
  c.geoid = geoid;
  c.geo_s0 = geo_s0;
  c.geo_s1 = geo_s1;
  c.ring_tab = ring_tab;
  c.rad_e = rad_e;
  MomSchedule* sc = nullptr;
  const int rc = moment_atm_launch(pl, c, lay, (char*)mom_ws, 0, 1, h0, h1, *slot_base, &sc, st);
  if (rc != MHB200_OK) return rc;
  *slot_base += sc->nslots;
  return MHB200_OK;
}

int moment_band_finish(mhb200_plan* pl, int nbands, int mode, int nslots, const double* dfactor, double* clm,
                       double* slm, void* mom_ws, cudaStream_t st) {
  const MomWs lay = mom_layout(pl, (size_t)pl->nlat + (size_t)nbands * pl->mom->ncta_max, g_units_bound(pl, 1));
  const int field_slot[2] = {0, nslots};
  const bool bc05 = (mode == MHB200_METHOD_BC05);
  return moment_reduce(pl, lay, (char*)mom_ws, 1, field_slot, bc05 ? 2 : 4, bc05 ? 2 : 1, dfactor, nullptr, clm, slm,
                       st);
}

bool moment_pressure_eligible(const mhb200_plan* pl, const double* plm_table) {
  return pl->mom != nullptr && plm_table == nullptr && moment_mode_env() >= 2;
}

// gen_pressure_stokes through K_A (pressure), K_B, K_C; sub-batches of nb_sub(pl) fields
int moment_pressure_run(mhb200_plan* pl, const double* P, const long long Ps[3], const double* G,
                        const long long Gs[3], const double* R, const long long Rs[3], int nbatch, double rad_e,
                        const double* dfactor, const double* dfactor_dev, double* clm, double* slm, void* mom_ws,
                        const int** flag_out, cudaStream_t st) {
  MomentPlan* mp = pl->mom;
  const int nbmax = std::min(nbatch, nb_sub(pl));
  const MomWs lay = mom_layout(pl, (size_t)nbmax * pl->nlat + mp->ncta_max, g_units_bound(pl, nbmax));
  char* ws = (char*)mom_ws;
  MHB_CUDA(cudaMemsetAsync(ws + lay.flag, 0, sizeof(int), st));
  *flag_out = (const int*)(ws + lay.flag);
  const size_t nout = (size_t)(pl->lmax + 1) * (pl->mmax + 1);
  const size_t smem = (size_t)(PNR * PF_RING + 2 * PNR) * sizeof(double);
  MHB_CUDA(cudaFuncSetAttribute(pressure_moment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int t0 = 0; t0 < nbatch; t0 += nbmax) {
    const int nb = std::min(nbmax, nbatch - t0);
    PressMomParams pp;
    memset(&pp, 0, sizeof(pp));
    pp.nlat = pl->nlat;
    pp.nlon = pl->nlon;
    pp.nchunks = mp->nchunks32;
    pp.ngroups = (pl->nlat + PNR - 1) / PNR;
    pp.nmb = pl->nmb;
    pp.nks_total = mp->nchunks * 16;
    pp.Frows = pl->Mpad;
    pp.col_grid = mp->col_grid32;
    pp.P = P + (size_t)t0 * Ps[0];
    pp.G = G + (size_t)t0 * Gs[0];
    pp.R = R + (size_t)t0 * Rs[0];
    for (int i = 0; i < 3; ++i) {
      pp.Ps[i] = Ps[i];
      pp.Gs[i] = Gs[i];
      pp.Rs[i] = Rs[i];
    }
    pp.trig = mp->trig;
    pp.F = (double*)(ws + lay.F);
    pp.slot_r0 = (double*)(ws + lay.slot_r0);
    pp.slot_q = (double*)(ws + lay.slot_q);
    pp.slot_ring = (int*)(ws + lay.slot_ring);
    pp.flag = (int*)(ws + lay.flag);
    pp.rad_e = rad_e;
    pp.inv_rad = 1.0 / rad_e;
    const double bound = MOM_ND_BOUND / (double)(pl->lmax + 2);
    unsigned long long bits;
    memcpy(&bits, &bound, 8);
    pp.dbound_hi = (unsigned)(bits >> 32);
    ring_center_kernel<<<dim3(pl->nlat, nb), 256, 0, st>>>(pp);
    g_launches.fetch_add(1);
    profile_begin(st);
    pressure_moment_kernel<<<dim3(pp.ngroups * pp.nmb, nb), 256, smem, st>>>(pp);
    profile_end(st);
    g_launches.fetch_add(1);
    MHB_CUDA(cudaGetLastError());
    int field_slot[NB_SUB + 1];
    for (int i = 0; i <= nb; ++i) field_slot[i] = i * pl->nlat;
    const int rc = moment_reduce(pl, lay, ws, nb, field_slot, 2, 0, dfactor, dfactor_dev, clm + t0 * nout,
                                 slm + t0 * nout, st);
    if (rc != MHB200_OK) return rc;
  }
  return MHB200_OK;
}

}  // namespace mhb

extern "C" {
// CPU-testable probe of the atmosphere kernel's chunk layout (no CUDA calls): sets holds 4 ints per contraction
// column (grid columns at phi, -phi, pi - phi, pi + phi: the output of mhb200_fold_probe at fold level 2).
// Fills col_grid[nchunks * 256] (-1 = inactive thread column), box_start[nchunks * 4], dir[4]; returns 1 when the
// sub-blocks are contiguous runs (TMA feed possible), 0 when not, -1 on bad arguments.
int mhb200_moment_layout_probe(int nf, const int* sets, int* col_grid, int* box_start, int* dir, int max_chunks) {
  if (nf < 1 || !sets || !col_grid || !box_start || !dir) return -1;
  const int nchunks = (nf + mhb::MKCF - 1) / mhb::MKCF;
  if (nchunks > max_chunks) return -1;
  std::vector<int> s4[4];
  for (int i = 0; i < nf; ++i)
    for (int k = 0; k < 4; ++k) s4[k].push_back(sets[4 * i + k]);
  const std::vector<int>* ptr[4] = {&s4[0], &s4[1], &s4[2], &s4[3]};
  std::vector<int> cg, bs;
  const bool ok = mhb::moment_layout(nf, ptr, dir, cg, bs);
  for (size_t i = 0; i < cg.size(); ++i) col_grid[i] = cg[i];
  for (size_t i = 0; i < bs.size(); ++i) box_start[i] = bs[i];
  return ok ? 1 : 0;
}

// CPU-testable probe of the moment path's static schedule (no CUDA calls): fills cta_unit[ncta + 1],
// cta_slot[ncta], field_slot[nb + 1]; returns the number of CTAs (or -1), *nslots_out = slots.
int mhb200_moment_schedule_probe(int nb, int nrings, int nchunks, int max_ctas, int* cta_unit, int* cta_slot,
                                 int* field_slot, int buf_ctas, int* nslots_out) {
  if (nb < 1 || nrings < 1 || nchunks < 1 || max_ctas < 1 || !cta_unit || !cta_slot || !field_slot || !nslots_out)
    return -1;
  std::vector<int> cu, cs, fs;
  int nslots = 0;
  const int ncta = mhb::moment_schedule(nb, nrings, nchunks, max_ctas, cu, cs, fs, &nslots);
  if (ncta + 1 > buf_ctas) return -1;
  for (int i = 0; i <= ncta; ++i) cta_unit[i] = cu[i];
  for (int i = 0; i < ncta; ++i) cta_slot[i] = cs[i];
  for (int i = 0; i <= nb; ++i) field_slot[i] = fs[i];
  *nslots_out = nslots;
  return ncta;
}
}
