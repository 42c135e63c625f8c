// Scattered point pressures -> Stokes coefficients (disc-load approximation) on B200.
//
// Replaces model_harmonics/gen_point_pressure.py:126-154 (disc Legendre recursion, radius
// powers) and :160-194 (_complex_harmonics: per-degree Legendre table, e^{i m phi} table,
// kron, 3-operand einsum) of the reference, including the third-party
// gravity_toolkit.legendre call at :184.
//
//   C_lm + i S_lm = dfactor_l * sum_p W_l[p] * Plm(cos th_p) * e^{i m phi_p}
//   W_l[p] = Pdisc_l(alpha_p) * (P_p/G_p) * (R_p/rad_e)^(l+2)
//
// Two paths, each ending in the same fixed-order reduce (points_reduce_kernel; deterministic, no fp64 atomics):
//
//  A. points_fused_kernel (the default while two of its CTAs fit on an SM: LMAX <= 61): one launch from the six
//     point arrays to the per-CTA partial sums -- see the comment above that kernel.
//
//  B. LMAX beyond that: points_prepare_kernel + points_contract_staged / _generic (round 1):
//  (1) weights role: one thread per point walks the degrees.  The disc factor
//      Pdisc_l = (P_{l-1}(alpha) - P_{l+1}(alpha))/2 is a cancellation-prone 3-term recursion
//      (error amplified by ~1/(1-alpha) ~ npts/2), so it is evaluated in EXACTLY the
//      reference's operation order with round-to-nearest intrinsics and no FMA contraction
//      (the degree-only quotients (2l+1)/(l+1), l/(l+1) come from a host table: IEEE division
//      gives the same bits on both sides).  The fused kernel walks the same recursion.
//  (2) trig role: the per (point, order) factors (u^m/1e-280) e^{i m phi}, eight orders per thread.
//  (3) points_contract_kernel: scatter-reduce.  A lane group of 4 lanes owns one pair of
//      orders (m, mtop-m) -- pairing gives every group the same number of degree steps -- and
//      each lane carries 8 points at a time through the Holmes-Featherstone recursion in
//      registers (Plm normalisation identical to legendre(NORMALIZE=True)).  All lanes of a warp
//      run ONE loop over the step index in lock step (a group changes from its first to its
//      second order at its own step, by reloading its state), so the cross-lane sums are
//      full-warp shuffles in uniform control flow.  Sums go to accumulators private to the lane
//      group in shared memory; per-CTA partials go to slots that the final kernel reduces.
#include "../../include/mhb200.h"
#include "common.cuh"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

namespace mhb {

constexpr int PT_NT = 128;       // threads per CTA (two CTAs per SM)
constexpr int PT_PER_LANE = 8;   // points carried per lane
constexpr int PT_GROUPS = 4;     // lanes per order pair
constexpr int PT_BATCH = PT_PER_LANE * PT_GROUPS;   // 32 points per pass
constexpr int PT_PAIRS = 32;     // order pairs per CTA (8 per warp)
constexpr int PT_MTILE = 2 * PT_PAIRS;

struct PointParams {
  long long npts, npad;
  int nbatch, lmax, mmax;
  const double *P, *G, *R, *phi, *theta, *area;
  long long Gs, Rs;
  double rad_e, area_denom;     // 2*pi*rad_e^2 formed on the host as the reference does
  double* W;                    // [nbatch][lmax+1][npad]
  double* xg;                   // [npad] cos(theta)
  double* ug;                   // [npad] sin(theta) (eps where 0), the Holmes-Featherstone scale base
  double2* E;                   // [mmax+1][npad] (u^m/1e-280) * (cos m phi, sin m phi)
  // contraction
  const double *f1, *f2, *pmm;  // Holmes-Featherstone multipliers [(lmax+1)][Mpad], sectoral seeds [Mpad]
  const double *dc1, *dc2;      // disc recursion quotients (2l+1)/(l+1), l/(l+1)  [lmax+1]
  int Mpad;
  int batches_per_cta;
  double* slots;                // [ncta_p][nbatch][2][lmax+1][mmax+1]
  int ncta_p;
  double *clm, *slm;
  // fused kernel
  const double* gl;             // [(lmax+1)][Mpad] multipliers of the one-multiplier recursion (points_fused_kernel)
  const double* scale;          // [(lmax+1)][Mpad] its normalisation c_lm, applied by the reduce kernel (or null)
  int tiles_per_cta, tiles_rem, fused_steps;   // CTA b takes tiles_per_cta (+1 if b < tiles_rem) consecutive tiles
  unsigned long long* flag;     // set to `gen` by the fast kernel when an angle is outside its short sincos
  unsigned long long gen;       // (a fresh value per call: the word never has to be cleared)
};

constexpr int DF_MAX = 448;      // degree factors travel as kernel parameters up to this many
struct DegreeFactors {
  const double* dev;             // used when lmax+1 > DF_MAX
  double v[DF_MAX];
};

// One launch prepares both operands of the contraction; blockIdx.y selects the role.
//   y <  nmg : per (point, order) factors for the orders [8 y, 8 y + 8)
//              E[m][p] = (u_p^m / 1e-280) * (cos(m phi_p), sin(m phi_p)),  u = sqrt(1 - x^2) (0 -> eps)
//              i.e. e^{i m phi} of gen_point_pressure.py:186-188 with the rescale of the scaled
//              Holmes-Featherstone column folded in; cos(theta), u and the power u^(8 y) are formed once per
//              thread and the power advanced by one multiplication per order
//   y >= nmg : field t = y - nmg: W_l[p] of gen_point_pressure.py:126-147 (and, for t = 0, cos(theta) and u)
constexpr int PT_MGROUP = 8;
__global__ void points_prepare_kernel(const PointParams q, int nmg) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= q.npad) return;
  if ((int)blockIdx.y < nmg) {
    const int m0 = blockIdx.y * PT_MGROUP;
    double u = 1.0, ph = 0.0, r = 0.0;
    const bool live = (p < q.npts);
    if (live) {
      const double x = cos(q.theta[p]);
      u = sqrt(1.0 - x * x);
      if (u == 0.0) u = 2.220446049250313e-16;
      ph = q.phi[p];
      // u^m0 / 1e-280 (1 for the zonal order): restarted from a short power at every group of orders
      r = (m0 == 0) ? 1.0 : 1.0e280;
      double bb = u;
      int k = m0;
      while (k) {
        if (k & 1) r *= bb;
        bb *= bb;
        k >>= 1;
      }
    }
#pragma unroll
    for (int i = 0; i < PT_MGROUP; ++i) {
      const int m = m0 + i;
      if (m > q.mmax) break;
      double2 e = make_double2(0.0, 0.0);
      if (live) {
        double sv, cv;
        sincos((double)m * ph, &sv, &cv);
        e = make_double2(r * cv, r * sv);
        r = (m == 0) ? 1.0e280 * u : r * u;
      }
      q.E[(size_t)m * q.npad + p] = e;
    }
    return;
  }
  const int t = blockIdx.y - nmg;
  double* W = q.W + (size_t)t * (q.lmax + 1) * q.npad + p;
  if (p >= q.npts) {
    for (int l = 0; l <= q.lmax; ++l) W[(size_t)l * q.npad] = 0.0;
    if (t == 0) {
      q.xg[p] = 0.0;
      q.ug[p] = 1.0;
    }
    return;
  }
  if (t == 0) {
    const double x = cos(q.theta[p]);
    double u = sqrt(1.0 - x * x);
    if (u == 0.0) u = 2.220446049250313e-16;
    q.xg[p] = x;
    q.ug[p] = u;
  }
  const double alpha = __dsub_rn(1.0, __ddiv_rn(q.area[p], q.area_denom));
  const double PG = __ddiv_rn(q.P[(size_t)t * q.npts + p], q.G[p * q.Gs]);
  const double r = __ddiv_rn(q.R[p * q.Rs], q.rad_e);
  double Pm1 = 1.0, Pl = 1.0;
  double pw = __dmul_rn(r, r);
  for (int l = 0; l <= q.lmax; ++l) {
    const double c1 = __ldg(q.dc1 + l), c2 = __ldg(q.dc2 + l);
    const double Pp1 = __dsub_rn(__dmul_rn(__dmul_rn(c1, alpha), Pl), __dmul_rn(c2, Pm1));
    const double Pdisc = __dmul_rn(__dsub_rn(Pm1, Pp1), 0.5);
    W[(size_t)l * q.npad] = __dmul_rn(__dmul_rn(Pdisc, PG), pw);
    pw = __dmul_rn(pw, r);
    Pm1 = Pl;
    Pl = Pp1;
  }
}

// grid: (ncta_p, n_mtiles, nbatch).  Fallback for degrees whose staged tables do not fit in shared
// memory (see points_contract_staged): W rows and multipliers come from global memory.
//
// Step s of a lane group with orders (ma, mb): s < sa = L+1-ma is degree ma+s of order ma, the
// rest is degree mb+(s-sa) of order mb; ma + mb is the same for every pair of the CTA's order
// tile, so every group takes nsteps = 2(L+1) - (ma+mb) steps.  The multiplier table makes the
// recursion uniform from the first degree: row l = m holds (0, -1) and row l = m+1 holds
// (sqrt(2m+3), 0), so with the state (p1, p2) = (0, pmm) the generic step
//   p_l = (x*f1)*p1 - f2*p2
// yields pmm at degree m and (x*sqrt(2m+3))*pmm at degree m+1 with the same roundings as the
// explicit seeds.
__global__ void __launch_bounds__(PT_NT, 2) points_contract_generic(const PointParams q) {
  extern __shared__ double acc_s[];   // [PT_PAIRS][2][pitch]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane & (PT_GROUPS - 1);
  const int pair = warp * (32 / PT_GROUPS) + lane / PT_GROUPS;
  const int t = blockIdx.z;
  const int L = q.lmax;
  // orders of this CTA's tile and of this lane group's pair
  const int m0 = blockIdx.y * PT_MTILE;
  const int m1 = min(m0 + PT_MTILE - 1, q.mmax);
  const int ma = m0 + pair, mb = m1 - pair;
  const bool has_a = (ma <= mb);          // pair exists
  const bool has_b = (ma < mb);           // second order distinct
  const int nsteps = 2 * (L + 1) - (m0 + m1);
  const int pitch = nsteps + 1;
  // groups without a pair walk order m0 with zero trig factors (they must stay in the warp's
  // shuffles); a group whose two orders coincide goes silent after its first walk
  const int mA = has_a ? ma : m0;
  const int mB = has_b ? mb : mA;
  const int sa = has_a ? (L + 1 - ma) : (nsteps + 4);
  double* acc_c = acc_s + (size_t)pair * 2 * pitch;
  // after the transposed reduction lane g of a group holds: 0 cos(s), 1 sin(s), 2 cos(s+1), 3 sin(s+1)
  double* acc_mine = acc_c + (g & 1) * pitch + (g >> 1);
  for (int i = tid; i < PT_PAIRS * 2 * pitch; i += PT_NT) acc_s[i] = 0.0;
  __syncthreads();

  const double* Wt = q.W + (size_t)t * (L + 1) * q.npad;
  const long long batch0 = (long long)blockIdx.x * q.batches_per_cta;
  const long long nbatches = (q.npad + PT_BATCH - 1) / PT_BATCH;
  const long long batch1 = min(batch0 + (long long)q.batches_per_cta, nbatches);
  const double pmmA = q.pmm[mA], pmmB = q.pmm[mB];

  struct RowSet {
    double4 w0, w1;
    double a, b;
  };

  for (long long b = batch0; b < batch1; ++b) {
    const long long pbase = b * PT_BATCH + g * PT_PER_LANE;       // npad is a multiple of PT_BATCH
    const double* Wb = Wt + pbase;
    double x[PT_PER_LANE], cm[PT_PER_LANE], sm[PT_PER_LANE], p1[PT_PER_LANE], p2[PT_PER_LANE];
    {
      const double4 xa = *reinterpret_cast<const double4*>(q.xg + pbase);
      const double4 xb = *reinterpret_cast<const double4*>(q.xg + pbase + 4);
      x[0] = xa.x; x[1] = xa.y; x[2] = xa.z; x[3] = xa.w;
      x[4] = xb.x; x[5] = xb.y; x[6] = xb.z; x[7] = xb.w;
    }
    auto load_state = [&](int m, bool live, double pmm) {
      const double2* Em = q.E + (size_t)m * q.npad + pbase;
#pragma unroll
      for (int i = 0; i < PT_PER_LANE; ++i) {
        const double2 e = live ? __ldg(Em + i) : make_double2(0.0, 0.0);
        cm[i] = e.x;
        sm[i] = e.y;
        p2[i] = pmm;
        p1[i] = 0.0;
      }
    };
    // row of step s: weights of the 8 points at that degree and the two recursion multipliers
    auto load_row = [&](RowSet& r, int s) {
      const bool first = (s < sa);
      const int m = first ? mA : mB;
      const int l = min(m + (first ? s : s - sa), L);
      const double* Wn = Wb + (size_t)l * q.npad;
      r.w0 = *reinterpret_cast<const double4*>(Wn);
      r.w1 = *reinterpret_cast<const double4*>(Wn + 4);
      r.a = __ldg(q.f1 + (size_t)l * q.Mpad + m);
      r.b = __ldg(q.f2 + (size_t)l * q.Mpad + m);
      asm volatile("prefetch.global.L1 [%0];" ::"l"(Wb + (size_t)min(l + 4, L) * q.npad));
    };
    // one degree step: advance the recursion and form this lane's partial sums over its 8 points
    auto step = [&](const RowSet& r, double& c, double& sn) {
      const double w[PT_PER_LANE] = {r.w0.x, r.w0.y, r.w0.z, r.w0.w, r.w1.x, r.w1.y, r.w1.z, r.w1.w};
      double c0 = 0.0, c1 = 0.0, s0 = 0.0, s1 = 0.0;   // two interleaved chains per component
#pragma unroll
      for (int i = 0; i < PT_PER_LANE; i += 2) {
        const double pa = fma(x[i] * r.a, p1[i], -(r.b * p2[i]));
        const double pb = fma(x[i + 1] * r.a, p1[i + 1], -(r.b * p2[i + 1]));
        p2[i] = p1[i];
        p1[i] = pa;
        p2[i + 1] = p1[i + 1];
        p1[i + 1] = pb;
        const double t0 = pa * w[i], t1 = pb * w[i + 1];
        c0 = fma(t0, cm[i], c0);
        c1 = fma(t1, cm[i + 1], c1);
        s0 = fma(t0, sm[i], s0);
        s1 = fma(t1, sm[i + 1], s1);
      }
      c = c0 + c1;
      sn = s0 + s1;
    };

    if (!has_a) {
#pragma unroll
      for (int i = 0; i < PT_PER_LANE; ++i) x[i] = 0.0;     // keeps the idle recursion bounded
    }
    load_state(mA, has_a, pmmA);
    if (b + 1 < batch1) {
      // next batch's first lines into L1 while this one runs
      asm volatile("prefetch.global.L1 [%0];" ::"l"(q.xg + pbase + PT_BATCH));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(q.E + (size_t)mA * q.npad + pbase + PT_BATCH));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(Wb + PT_BATCH + (size_t)mA * q.npad));
    }
    asm volatile("prefetch.global.L1 [%0];" ::"l"(q.E + (size_t)mB * q.npad + pbase));
    RowSet ra, rb;
    load_row(ra, 0);
    load_row(rb, 1);
    for (int s = 0; s < nsteps; s += 2) {
      double cA, sA, cB, sB;
      if (s == sa) load_state(mB, has_b, pmmB);
      step(ra, cA, sA);
      load_row(ra, s + 2);
      if (s + 1 == sa) load_state(mB, has_b, pmmB);
      step(rb, cB, sB);
      load_row(rb, s + 3);
      // transposed reduction over the 4 lanes of the group: 3 exchanges for the 4 sums
      const bool hi = (g & 2) != 0, odd = (g & 1) != 0;
      const double k0 = (hi ? cB : cA) + __shfl_xor_sync(0xffffffffu, hi ? cA : cB, 2);
      const double k1 = (hi ? sB : sA) + __shfl_xor_sync(0xffffffffu, hi ? sA : sB, 2);
      const double v = (odd ? k1 : k0) + __shfl_xor_sync(0xffffffffu, odd ? k0 : k1, 1);
      acc_mine[s] += v;          // s + 1 == nsteps lands in the spare entry pitch - 1
    }
  }
  __syncthreads();
  // write this CTA's partial sums
  double* slot = q.slots + ((size_t)blockIdx.x * q.nbatch + t) * 2 * (L + 1) * (q.mmax + 1);
  const size_t plane = (size_t)(L + 1) * (q.mmax + 1);
  for (int pr = warp; pr < PT_PAIRS; pr += PT_NT / 32) {
    const int a = m0 + pr, bo = m1 - pr;
    if (a > bo) continue;
    const double* ac = acc_s + (size_t)pr * 2 * pitch;
    const int na = L + 1 - a;
    const int nb = (a < bo) ? (L + 1 - bo) : 0;
    for (int i = lane; i < na + nb; i += 32) {
      const int m = (i < na) ? a : bo;
      const int l = (i < na) ? (a + i) : (bo + i - na);
      slot[(size_t)l * (q.mmax + 1) + m] = ac[i];
      slot[plane + (size_t)l * (q.mmax + 1) + m] = ac[pitch + i];
    }
  }
}

// ---------------------------------------------------------------------------------------
// Staged variant (used whenever its shared memory fits): nothing in the degree loop touches
// global memory.  Same lane-group shape as above (4 lanes per order pair, 8 points per lane; an
// 8-lane x 4-point shape with twice the warps was measured slower: the cross-lane reduction and
// the per-step fixed work grow faster than the latency hiding improves).
//   fab_s   per lane group, the recursion multipliers of its whole walk indexed by step
//           (they depend on the group's orders only: built once per CTA, reused by every batch)
//   wt_s    the W rows of the batch's 32 points for all degrees plus their cos(theta) row,
//           double-buffered: batch b+1 is copied (cp.async) while batch b is contracted
//   eb_s    each lane's own 8 trig factors: second order of the current batch (copied at the start
//           of the batch, read at the group's switch step), then first order of the next batch
//           (copied at the switch step, read at the start of the next batch).  Producer =
//           consumer, so cp.async.wait_group is the only synchronisation.
// A W row is stored as [half][lane][4 points] with a 16-byte pad, so the 128-bit reads of a
// quarter warp (two lane groups at consecutive degrees) fall into distinct banks.
// ---------------------------------------------------------------------------------------
constexpr int WT_PITCH = 34;     // doubles per staged row

__host__ __device__ inline int staged_fpitch(int nsteps) { return (nsteps + 2) | 1; }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc));
}

__global__ void __launch_bounds__(PT_NT, 2) points_contract_staged(const PointParams q) {
  extern __shared__ __align__(16) double acc_s[];   // [PT_PAIRS][2][pitch] | fab_s | wt_s | eb_s
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane & (PT_GROUPS - 1);
  const int pair = warp * (32 / PT_GROUPS) + lane / PT_GROUPS;
  const int t = blockIdx.z;
  const int L = q.lmax;
  const int m0 = blockIdx.y * PT_MTILE;
  const int m1 = min(m0 + PT_MTILE - 1, q.mmax);
  const int ma = m0 + pair, mb = m1 - pair;
  const bool has_a = (ma <= mb);
  const bool has_b = (ma < mb);
  const int nsteps = 2 * (L + 1) - (m0 + m1);
  const int pitch = nsteps + 1;
  const int fpitch = staged_fpitch(nsteps);
  const int nrows = L + 2;                            // degrees 0..L and the cos(theta) row
  const int mA = has_a ? ma : m0;
  const int mB = has_b ? mb : mA;
  const int sa = has_a ? (L + 1 - ma) : (nsteps + 4);
  double* acc_c = acc_s + (size_t)pair * 2 * pitch;
  // after the transposed reduction lane g of a group holds: 0 cos(s), 1 sin(s), 2 cos(s+1), 3 sin(s+1)
  double* acc_mine = acc_c + (g & 1) * pitch + (g >> 1);
  double2* fab_s = reinterpret_cast<double2*>(acc_s + (size_t)PT_PAIRS * 2 * pitch);
  double* wt_s = reinterpret_cast<double*>(fab_s + (size_t)PT_PAIRS * fpitch);
  double2* eb_mine = reinterpret_cast<double2*>(wt_s + (size_t)2 * nrows * WT_PITCH) + tid * PT_PER_LANE;
  const double2* fab_mine = fab_s + (size_t)pair * fpitch;

  const double* Wt = q.W + (size_t)t * (L + 1) * q.npad;
  const long long batch0 = (long long)blockIdx.x * q.batches_per_cta;
  const long long nbatches = (q.npad + PT_BATCH - 1) / PT_BATCH;
  const long long batch1 = min(batch0 + (long long)q.batches_per_cta, nbatches);

  // copy of one batch's rows: 16-byte pieces, piece ch of a row = points 2ch, 2ch+1
  auto issue_tile = [&](long long bb, int buf) {
    const double* src = Wt + bb * PT_BATCH;
    const double* xsrc = q.xg + bb * PT_BATCH;
    double* dst = wt_s + (size_t)buf * nrows * WT_PITCH;
    for (int c = tid; c < nrows * 16; c += PT_NT) {
      const int row = c >> 4, ch = c & 15;
      const double* gsrc = ((row <= L) ? src + (size_t)row * q.npad : xsrc) + 2 * ch;
      cp_async16(dst + row * WT_PITCH + ((ch >> 1) & 1) * 16 + (ch >> 2) * 4 + (ch & 1) * 2, gsrc);
    }
  };
  if (batch0 < batch1) issue_tile(batch0, 0);
  asm volatile("cp.async.commit_group;");

  for (int i = tid; i < PT_PAIRS * 2 * pitch; i += PT_NT) acc_s[i] = 0.0;
  // (blocks of 8 entries per thread: the table loads of a block are independent, so their latencies overlap)
  for (int i0 = tid; i0 < PT_PAIRS * fpitch; i0 += 8 * PT_NT) {
    double2 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = i0 + k * PT_NT;
      const int pr = i / fpitch, s = i - pr * fpitch;
      const int a = m0 + pr, bo = m1 - pr, na = L + 1 - a;
      // idle groups, silent second walks, the step past the end: multipliers 0, so p stays 0
      int l = 0, m = 0;
      bool live = (i < PT_PAIRS * fpitch) && (a <= bo) && (s < nsteps);
      if (s < na) {
        l = a + s;
        m = a;
      } else if (a < bo) {
        l = bo + (s - na);
        m = bo;
      } else {
        live = false;
      }
      const size_t o = live ? (size_t)l * q.Mpad + m : 0;
      const double f1v = __ldg(q.f1 + o), f2v = __ldg(q.f2 + o);
      v[k] = live ? make_double2(f1v, f2v) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = i0 + k * PT_NT;
      if (i < PT_PAIRS * fpitch) fab_s[i] = v[k];
    }
  }
  const double pmmA = q.pmm[mA], pmmB = q.pmm[mB];
  const int winc = has_a ? WT_PITCH : 0;              // idle groups stay on one row
  const bool hi = (g & 2) != 0, odd = (g & 1) != 0;

  for (long long b = batch0; b < batch1; ++b) {
    const int buf = (int)(b - batch0) & 1;
    const long long pbase = b * PT_BATCH + g * PT_PER_LANE;
    // this batch's rows (and this lane's first-order trig factors) have landed, and every warp is
    // done with the previous batch
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const double* wb = wt_s + (size_t)buf * nrows * WT_PITCH + g * 4;
    double x[PT_PER_LANE], cm[PT_PER_LANE], sm[PT_PER_LANE], p1[PT_PER_LANE], p2[PT_PER_LANE];
    {
      const double4 xa = *reinterpret_cast<const double4*>(wb + (L + 1) * WT_PITCH);
      const double4 xb = *reinterpret_cast<const double4*>(wb + (L + 1) * WT_PITCH + 16);
      x[0] = xa.x; x[1] = xa.y; x[2] = xa.z; x[3] = xa.w;
      x[4] = xb.x; x[5] = xb.y; x[6] = xb.z; x[7] = xb.w;
    }
    {
      // first order: the lane's own staged copy (made at the previous batch's switch step); the
      // CTA's first batch reads global memory
      const double2* Ea = q.E + (size_t)mA * q.npad + pbase;
      const bool staged = (b != batch0);
#pragma unroll
      for (int i = 0; i < PT_PER_LANE; ++i) {
        const double2 e = has_a ? (staged ? eb_mine[i] : __ldg(Ea + i)) : make_double2(0.0, 0.0);
        cm[i] = e.x;
        sm[i] = e.y;
        p2[i] = pmmA;
        p1[i] = 0.0;
      }
    }
    // group 1: second-order trig factors of this batch into the lane's slot (just read above)
    if (has_b) {
      const double2* Eb = q.E + (size_t)mB * q.npad + pbase;
#pragma unroll
      for (int i = 0; i < PT_PER_LANE; ++i) cp_async16(eb_mine + i, Eb + i);
    }
    asm volatile("cp.async.commit_group;");
    // group 2: the next batch's rows into the other buffer
    const bool more = (b + 1 < batch1);
    if (more) {
      issue_tile(b + 1, buf ^ 1);
      if (has_a) asm volatile("prefetch.global.L2 [%0];" ::"l"(q.E + (size_t)mA * q.npad + pbase + PT_BATCH));
    }
    asm volatile("cp.async.commit_group;");

    // wrow walks the staged rows one degree per step (the step past the last degree reads the
    // cos(theta) row: finite values that end in the spare accumulator entry)
    const double* wrow = wb + mA * WT_PITCH;
    // change to the group's second order: trig factors from the lane's slot, which is then refilled
    // with the first-order factors of the next batch (group 3)
    auto switch_state = [&]() {
      asm volatile("cp.async.wait_group 1;" ::: "memory");
#pragma unroll
      for (int i = 0; i < PT_PER_LANE; ++i) {
        const double2 e = has_b ? eb_mine[i] : make_double2(0.0, 0.0);
        cm[i] = e.x;
        sm[i] = e.y;
        p2[i] = pmmB;
        p1[i] = 0.0;
      }
      wrow = wb + mB * WT_PITCH;
      if (more) {
        const double2* En = q.E + (size_t)mA * q.npad + pbase + PT_BATCH;
#pragma unroll
        for (int i = 0; i < PT_PER_LANE; ++i) cp_async16(eb_mine + i, En + i);
      }
      asm volatile("cp.async.commit_group;");
    };
    auto step = [&](const double2 ab, double& c, double& sn) {
      const double4 w0 = *reinterpret_cast<const double4*>(wrow);
      const double4 w1 = *reinterpret_cast<const double4*>(wrow + 16);
      wrow += winc;
      const double w[PT_PER_LANE] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      double c0 = 0.0, c1 = 0.0, s0 = 0.0, s1 = 0.0;   // two interleaved chains per component
#pragma unroll
      for (int i = 0; i < PT_PER_LANE; i += 2) {
        const double pa = fma(x[i] * ab.x, p1[i], -(ab.y * p2[i]));
        const double pb = fma(x[i + 1] * ab.x, p1[i + 1], -(ab.y * p2[i + 1]));
        p2[i] = p1[i];
        p1[i] = pa;
        p2[i + 1] = p1[i + 1];
        p1[i + 1] = pb;
        const double t0 = pa * w[i], t1 = pb * w[i + 1];
        c0 = fma(t0, cm[i], c0);
        c1 = fma(t1, cm[i + 1], c1);
        s0 = fma(t0, sm[i], s0);
        s1 = fma(t1, sm[i + 1], s1);
      }
      c = c0 + c1;
      sn = s0 + s1;
    };
    // The 4 lanes of a group are combined by a transposed reduction (3 exchanges for the 4 sums of
    // two steps).  It runs one iteration late, its two exchange stages placed inside the next two
    // steps, so that the shuffle latency overlaps their arithmetic.
    double qcA = 0.0, qsA = 0.0, qcB = 0.0, qsB = 0.0;
    int qs = 0;
    double2 abA = fab_mine[0], abB = fab_mine[1];
    for (int s = 0; s < nsteps; s += 2) {
      double cA, sA, cB, sB;
      const double2 nA = fab_mine[s + 2], nB = fab_mine[s + 3];     // past the end: unused values
      if (s == sa) switch_state();
      const double r0 = __shfl_xor_sync(0xffffffffu, hi ? qcA : qcB, 2);
      const double r1 = __shfl_xor_sync(0xffffffffu, hi ? qsA : qsB, 2);
      step(abA, cA, sA);
      const double k0 = (hi ? qcB : qcA) + r0;
      const double k1 = (hi ? qsB : qsA) + r1;
      if (s + 1 == sa) switch_state();
      const double r2 = __shfl_xor_sync(0xffffffffu, odd ? k0 : k1, 1);
      step(abB, cB, sB);
      acc_mine[qs] += (odd ? k1 : k0) + r2;
      qcA = cA; qsA = sA; qcB = cB; qsB = sB;
      qs = s;
      abA = nA; abB = nB;
    }
    {
      const double k0 = (hi ? qcB : qcA) + __shfl_xor_sync(0xffffffffu, hi ? qcA : qcB, 2);
      const double k1 = (hi ? qsB : qsA) + __shfl_xor_sync(0xffffffffu, hi ? qsA : qsB, 2);
      acc_mine[qs] += (odd ? k1 : k0) + __shfl_xor_sync(0xffffffffu, odd ? k0 : k1, 1);   // qs + 1 == nsteps
    }                                                                                      // lands in the spare entry
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // write this CTA's partial sums
  double* slot = q.slots + ((size_t)blockIdx.x * q.nbatch + t) * 2 * (L + 1) * (q.mmax + 1);
  const size_t plane = (size_t)(L + 1) * (q.mmax + 1);
  for (int pr = warp; pr < PT_PAIRS; pr += PT_NT / 32) {
    const int a = m0 + pr, bo = m1 - pr;
    if (a > bo) continue;
    const double* ac = acc_s + (size_t)pr * 2 * pitch;
    const int na = L + 1 - a;
    const int nb = (a < bo) ? (L + 1 - bo) : 0;
    for (int i = lane; i < na + nb; i += 32) {
      const int m = (i < na) ? a : bo;
      const int l = (i < na) ? (a + i) : (bo + i - na);
      slot[(size_t)l * (q.mmax + 1) + m] = ac[i];
      slot[plane + (size_t)l * (q.mmax + 1) + m] = ac[pitch + i];
    }
  }
}

// ---------------------------------------------------------------------------------------
// Fused variant (the default whenever two of its CTAs fit on an SM: LMAX <= 61, the reference's default 60
// included).  One launch reads the six point arrays and writes the per-CTA partial sums; the weights W_l[p] and
// the factors u^m e^{i m phi} never travel through HBM (C2: 152 MB of DRAM traffic -> 5 MB, 0.214 -> 0.143 ms).
//
//  * Recursion with ONE multiplier.  With q_l = Plm / (u^m c_lm), c_lm = f2_lm c_(l-2)m, the
//    Holmes-Featherstone step p_l = x f1 p_(l-1) - f2 p_(l-2) becomes q_l = (x g_lm) q_(l-1) - q_(l-2),
//    g_lm = f1_lm c_(l-1)m / c_lm (host table, formed in extended precision): 5 fp64 operations per
//    (point, l, m) instead of 6, and the normalisation c_lm leaves the point sum (points_reduce_kernel).
//  * Two orders per lane at the same degree (a "duo" m, m+1), so one shared-memory read of W_l[p] serves two
//    (point, l, m) elements: the round-1 kernel read 8 bytes per element and sat at ~90 % of the
//    shared-memory bandwidth next to its arithmetic.
//  * A warp = 4 duos (8 consecutive orders) x 8 lanes x 8 points per lane = one 64-point tile.  All lanes of
//    a warp are at the SAME degree: the walk runs from the first order of the 8 to LMAX, an order m that
//    starts d steps later idles through d steps with multiplier 0 (its state just rotates through
//    0, +-1; those sums land in (l < m) entries that are never written out) and the state is seeded so
//    that it arrives at (q_(m-1), q_m) = (0, 1) on time.  Warp w walks orders [8w, 8w+8) and then
//    [8(7-w), 8(7-w)+8) of the CTA's 64-order tile: equal work for every warp, the change of orders is
//    uniform across the warp (the e^{i m phi} factors of the second walk are computed there, with no
//    divergence), and warps only meet at the tile boundaries.
//  * The 8 lanes of a duo are combined every two steps by a transposed reduction (7 exchanges for 8 sums, lane
//    roles chosen so that nothing has to be selected, run one iteration late under the next one's arithmetic);
//    each lane then owns one accumulator in shared memory.
//  * The next tile is prepared WHILE this one is contracted (two tile buffers) by four PRODUCER warps of the
//    same CTA: warps 4-5 walk the disc-weight recursion of one point per thread -- the reference's exact
//    operation order, as points_prepare_kernel -- and warps 6-7 write the per-point rows (cos theta, u, phi,
//    e^{i phi}, u^(8k)); both are latency-bound and short.  Named barriers hand the buffers over ("tile ready"
//    per contraction warp, "buffer free" per buffer), so the contraction warps never wait for one another.
//    The register file is split to match (setmaxnreg): 216 registers per contraction thread, 40 per producer.
//  * e^{i m phi} of the ROUNDED product m phi (what the reference's exp(1j m phi) sees) costs one short
//    sincos per (point, duo) -- Cody-Waite + Taylor, coefficients in the constant bank -- and one rotation by
//    e^{i phi} for the duo's second order, with the rounding difference of the two products applied exactly.
// ---------------------------------------------------------------------------------------
constexpr int PF_NT = 256;       // 4 contraction warps + 4 producer warps
constexpr int PF_REGS_MAIN = 216, PF_REGS_PROD = 40;   // 128 x (216 + 40) = 256 x 128 registers per CTA
constexpr int PF_P = 8;          // points per lane
constexpr int PF_TILE = 64;      // points per tile: 8 lanes x 8 points
constexpr int PF_ORD = 64;       // orders per CTA (4 warps x 2 walks x 8 orders)
constexpr int PF_AUX = 13;       // per-point rows next to the weights: x, phi, cos phi, sin phi, u, u^(m0 + 8k) k = 0..7

// rows of one warp's accumulators: both walks, each rounded up to an even number of steps
__host__ __device__ inline int fused_walk_steps(int lmax, int mmax, int m0, int quad) {
  const int mq = m0 + 8 * quad;
  const int n = (mq <= mmax) ? (lmax + 1 - mq) : 0;
  return (n + 1) & ~1;
}

__device__ __forceinline__ void pf_load(const double* row, double (&v)[PF_P]) {
#pragma unroll
  for (int k = 0; k < PF_P / 2; ++k) {
    const double2 a = *reinterpret_cast<const double2*>(row + 16 * k);
    v[2 * k] = a.x;
    v[2 * k + 1] = a.y;
  }
}

// sin and cos of a (|a| < 2^19: Cody-Waite reduction with a three-part pi/2 whose leading parts have 33
// bits, Taylor polynomials on [-pi/4, pi/4] whose truncation is < 3e-18; within 2 ulp of NumPy's on 2e6 arguments
// over the whole range, tests/test_gpu_edge_cases.py, like the library function).  ~25 fp64 operations and no
// conversion instructions, against ~115 instructions for the general-purpose sincos().  The coefficients sit
// in the constant bank, where the FMAs read them as operands.
__constant__ double PF_SC[20] = {
    0.6366197723675814, 6755399441055744.0, 1.5707963267341256, 6.077100506303966e-11, 2.0222662487959506e-21,
    // sin: 1/3!, -1/5!, ... -1/17!  (of r - r z (c0 + z (c1 + ...)))
    1.0 / 6.0, -1.0 / 120.0, 1.0 / 5040.0, -1.0 / 362880.0, 1.0 / 39916800.0, -1.0 / 6227020800.0,
    1.0 / 1307674368000.0, -1.0 / 355687428096000.0,
    // cos: -1/2!, 1/4!, ... 1/16!  (of 1 + z (c0 + z (c1 + ...)))
    -0.5, 1.0 / 24.0, -1.0 / 720.0, 1.0 / 40320.0, -1.0 / 3628800.0, 1.0 / 479001600.0, -1.0 / 87178291200.0};
__constant__ double PF_C16 = 1.0 / 20922789888000.0;
constexpr double PF_TRIG_MAX = 500000.0;      // beyond this the general-purpose functions take over

__device__ __forceinline__ void pf_sincos(const double a, double& sn, double& cs) {
  const double t = fma(a, PF_SC[0], PF_SC[1]);
  const int n = __double2loint(t);
  const double k = t - PF_SC[1];
  double r = fma(-k, PF_SC[2], a);
  r = fma(-k, PF_SC[3], r);
  r = fma(-k, PF_SC[4], r);
  const double z = r * r;
  double ps = PF_SC[12];
#pragma unroll
  for (int i = 11; i >= 5; --i) ps = fma(ps, z, PF_SC[i]);
  double pc = PF_C16;
#pragma unroll
  for (int i = 19; i >= 13; --i) pc = fma(pc, z, PF_SC[i]);
  const double s0 = fma(-(z * r), ps, r);
  const double c0 = fma(z, pc, 1.0);
  // quadrant: n & 1 swaps, bit 1 of n (sin) / of n + 1 (cos) negates
  const double sv = (n & 1) ? c0 : s0, cv = (n & 1) ? s0 : c0;
  sn = (n & 2) ? -sv : sv;
  cs = ((n + 1) & 2) ? -cv : cv;
}

__global__ void sincos_probe_kernel(const double* a, long long n, double* sn, double* cs) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) pf_sincos(a[i], sn[i], cs[i]);
}

// state of one point's disc-weight recursion (gen_point_pressure.py:126-147), advanced one degree at a time
struct DiscChain {
  double alpha, PG, r, Pm1, Pl, pw;
  bool live;
};

__device__ __forceinline__ void chain_start(DiscChain& c, const PointParams& q, const int t, const long long p) {
  c.live = (p < q.npts);
  c.alpha = c.PG = c.r = 0.0;
  if (c.live) {
    c.alpha = __dsub_rn(1.0, __ddiv_rn(q.area[p], q.area_denom));
    c.PG = __ddiv_rn(q.P[(size_t)t * q.npts + p], q.G[p * q.Gs]);
    c.r = __ddiv_rn(q.R[p * q.Rs], q.rad_e);
  }
  c.Pm1 = 1.0;
  c.Pl = 1.0;
  c.pw = __dmul_rn(c.r, c.r);
}

// W_l of the chain's point (the reference's operation order, no FMA contraction); dc = ((2l+1)/(l+1), l/(l+1))
__device__ __forceinline__ double chain_step(DiscChain& c, const double2 dc) {
  const double Pp1 = __dsub_rn(__dmul_rn(__dmul_rn(dc.x, c.alpha), c.Pl), __dmul_rn(dc.y, c.Pm1));
  const double Pdisc = __dmul_rn(__dsub_rn(c.Pm1, Pp1), 0.5);
  const double w = __dmul_rn(__dmul_rn(Pdisc, c.PG), c.pw);
  c.pw = __dmul_rn(c.pw, c.r);
  c.Pm1 = c.Pl;
  c.Pl = Pp1;
  return c.live ? w : 0.0;
}

// per-point rows of one tile column: x, phi, cos phi, sin phi, u, u^(m0 + 8k)
template <bool SAFE>
__device__ __forceinline__ void aux_column(double* col, const PointParams& q, const long long p, const int m0) {
  double x = 0.0, u = 1.0, ph = 0.0, cp = 1.0, sp = 0.0;
  if (p < q.npts) {
    const double th = q.theta[p];
    ph = q.phi[p];
    if (SAFE) {
      x = cos(th);
      sincos(ph, &sp, &cp);
    } else {
      // an angle (or its largest multiple in this CTA) outside the short sincos: the safe kernel redoes the call
      if (!(fabs(th) < PF_TRIG_MAX) || !(fabs(ph) * (double)(m0 + PF_ORD) < PF_TRIG_MAX)) *q.flag = q.gen;
      double sth;
      pf_sincos(th, sth, x);
      pf_sincos(ph, sp, cp);
    }
    u = sqrt(1.0 - x * x);
    if (u == 0.0) u = 2.220446049250313e-16;
  }
  col[0 * PF_TILE] = x;
  col[1 * PF_TILE] = ph;
  col[2 * PF_TILE] = cp;
  col[3 * PF_TILE] = sp;
  col[4 * PF_TILE] = u;
  double pw = 1.0, bb = u;
  for (int k = m0; k; k >>= 1) {
    if (k & 1) pw *= bb;
    bb *= bb;
  }
  const double u2 = u * u, u4 = u2 * u2, u8 = u4 * u4;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    col[(5 + k) * PF_TILE] = pw;
    pw *= u8;
  }
}

__device__ __forceinline__ void pf_bar_sync(const int id, const int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ void pf_bar_arrive(const int id, const int n) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}
// named barriers: 1 producers only; 2 + 2 w + parity "tile ready" for contraction warp w (128 producer arrivals +
// the warp's 32); 10 + parity "buffer free" (128 contraction arrivals + the 128 producers)
constexpr int PF_BAR_PROD = 1, PF_BAR_FULL = 2, PF_BAR_EMPTY = 10;

// SAFE = false: the kernel as described.  SAFE = true: the same walk with the general-purpose cos / sincos for
// every angle, one CTA per SM and no register hand-over (ptxas 12.9 does not survive a device-function call
// under setmaxnreg); it is launched after the fast kernel on every call and returns at once unless that one met
// an angle beyond PF_TRIG_MAX (no real longitude does), in which case it overwrites all the partial sums.
template <bool SAFE>
__global__ void __launch_bounds__(PF_NT, SAFE ? 1 : 2) points_fused_kernel(const PointParams q) {
  constexpr int P = PF_P, TILE = PF_TILE;
  if (SAFE) {
    if (*q.flag != q.gen) return;
  }
  extern __shared__ __align__(16) double acc_s[];   // acc [4][NS][16] | dc [L+1][2] | 2 x tile [L+3+AUX][TILE]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = lane >> 3, j = lane & 7;
  const int t = blockIdx.z;
  const int L = q.lmax;
  const int m0 = blockIdx.y * PF_ORD;
  const int NS = q.fused_steps;
  double* acc_w = acc_s + (size_t)warp * NS * 16;
  double* dc_s = acc_s + (size_t)4 * NS * 16;
  double* tiles = dc_s + 2 * ((L + 2) & ~1);
  // rows 0 .. L: weights, L + 1: zeros, L + 2: scratch, then the per-point rows
  const size_t tile_doubles = (size_t)(L + 3 + PF_AUX) * TILE;
  const size_t aux_off = (size_t)(L + 3) * TILE;
  const int mq2[2] = {m0 + 8 * warp, m0 + 8 * (7 - warp)};
  const int n2_[2] = {fused_walk_steps(L, q.mmax, m0, warp), fused_walk_steps(L, q.mmax, m0, 7 - warp)};

  const long long tile0 = (long long)blockIdx.x * q.tiles_per_cta + min((int)blockIdx.x, q.tiles_rem);
  const long long tile1 = tile0 + q.tiles_per_cta + (((int)blockIdx.x < q.tiles_rem) ? 1 : 0);

  if (warp >= 4) {
    // ---- producers: tile tb + 1 into the other buffer while the contraction warps work on tile tb ----
    if (!SAFE) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PF_REGS_PROD));
    // the recursion's fp64 work loads the two SM sub-partitions its warps sit on: the second half of the grid (the
    // CTAs that usually share an SM with the first half) puts it on the other two
    const int pt = (tid - 128) ^ ((blockIdx.x * 2 >= gridDim.x) ? 64 : 0);
    const int pcol = pt & 63;
    for (int i = pt; i <= L; i += 128) {
      dc_s[2 * i] = __ldg(q.dc1 + i);
      dc_s[2 * i + 1] = __ldg(q.dc2 + i);
    }
    for (int i = pt; i < 2 * TILE; i += 128)           // the step past LMAX of an odd walk reads zeros
      tiles[(size_t)(i / TILE) * tile_doubles + (size_t)(L + 1) * TILE + (i % TILE)] = 0.0;
    pf_bar_sync(PF_BAR_PROD, 128);                     // dc_s is in place (producers only)
    for (long long tb = tile0; tb < tile1 + 2; ++tb) {
      const int par = (int)(tb - tile0) & 1;
      // the buffer is free once every contraction warp is done with tile tb - 2 (the passes past the last tile
      // only match those arrivals)
      if (tb >= tile0 + 2) pf_bar_sync(PF_BAR_EMPTY + par, 256);
      if (tb < tile1) {
        double* tn = tiles + (size_t)par * tile_doubles;
        const long long p = tb * TILE + pcol;
        if (pt < 64) {
          DiscChain ch;
          chain_start(ch, q, t, p);
          // blocks of four degrees: the loop-carried part (two dependent operations per degree) first, then
          // the four independent tails, so that their latencies overlap
          for (int l0 = 0; l0 <= L; l0 += 4) {
            double w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
              w[k] = chain_step(ch, *reinterpret_cast<const double2*>(dc_s + 2 * min(l0 + k, L)));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (l0 + k <= L) tn[(size_t)(l0 + k) * TILE + pcol] = w[k];
          }
        } else {
          aux_column<SAFE>(tn + aux_off + pcol, q, p, m0);
        }
#pragma unroll
        for (int w = 0; w < 4; ++w) pf_bar_arrive(PF_BAR_FULL + 2 * w + par, 160);     // tile tb is complete
        // the inputs of the tile after towards L2
        const long long pn = p + TILE;
        if (tb + 1 < tile1 && pn < q.npts) {
          if (pt < 64) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(q.area + pn));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(q.P + (size_t)t * q.npts + pn));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(q.G + pn * q.Gs));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(q.R + pn * q.Rs));
          } else {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(q.theta + pn));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(q.phi + pn));
          }
        }
      }
    }
    return;
  }

  // ---- contraction warps ----
  if (!SAFE) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(PF_REGS_MAIN));
  for (int i = lane; i < NS * 16; i += 32) acc_w[i] = 0.0;
  // Roles inside a duo's 8 lanes (bits of j): b2 picks which of the two orders is the lane's slot A (the other is
  // slot B), b1 which of cos / sin is the "first" component.  The lanes combine their sums by a transposed
  // reduction in which every lane KEEPS its slot-A / first sums and SENDS its slot-B / second sums: the partner
  // across b2 (b1) has the roles swapped, so no value has to be selected.  Only the last exchange (b0: which of
  // the two steps) selects.  The lane ends up with: order ma + b2, component b1, step s + b0.
  const bool b2 = (j & 4) != 0, b1 = (j & 2) != 0, b0 = (j & 1) != 0;
  const int acc_lane = (b0 ? 16 : 0) + grp * 4 + (b2 ? 2 : 0) + (b1 ? 1 : 0);

  for (long long tb = tile0; tb < tile1; ++tb) {
    const int cur = (int)(tb - tile0) & 1;
    pf_bar_sync(PF_BAR_FULL + 2 * warp + cur, 160);    // this tile's buffer is complete (the warps do not wait
                                                       // for one another: only the producers see all of them)
    const double* tile = tiles + (size_t)cur * tile_doubles;
    const double* aux = tile + aux_off;

    double x[P];
    pf_load(aux + 2 * j, x);

    // the two walks of this warp: orders [mq, mq + 8), degrees mq .. mq + n2 - 1
    int row0 = 0;
#pragma unroll 1
    for (int wk = 0; wk < 2; ++wk) {
      const int mq = mq2[wk], n2 = n2_[wk];
      if (n2 == 0) continue;
      const int quad = wk ? 7 - warp : warp;
      const int ma = mq + 2 * grp;               // this lane's orders: ma, ma + 1
      // slot A / B: first and second trig factor, recursion state
      double fa[P], ga[P], fb[P], gb[P], pa1[P], pa2[P], pb1[P], pb2[P];
      {
        const bool live_a = (ma <= q.mmax), live_b = (ma + 1 <= q.mmax);
        // seeds: an order that starts d steps into the walk (d = 2 grp for ma, 2 grp + 1 for ma + 1) arrives at
        // (q_(m-1), q_m) = (0, 1) on time
        const double sgn = (grp & 1) ? 1.0 : -1.0;
        const double dma = (double)ma, dmb = (double)(ma + 1);
        const double* ax = aux + 2 * j;
#pragma unroll
        for (int k = 0; k < P / 2; ++k) {
          const double2 ph2 = *reinterpret_cast<const double2*>(ax + 1 * TILE + 16 * k);
          const double2 cp2 = *reinterpret_cast<const double2*>(ax + 2 * TILE + 16 * k);
          const double2 sp2 = *reinterpret_cast<const double2*>(ax + 3 * TILE + 16 * k);
          const double2 u2_ = *reinterpret_cast<const double2*>(ax + 4 * TILE + 16 * k);
          const double2 um2 = *reinterpret_cast<const double2*>(ax + (size_t)(5 + quad) * TILE + 16 * k);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int i = 2 * k + h;
            const double ph = h ? ph2.y : ph2.x, cp = h ? cp2.y : cp2.x, sp = h ? sp2.y : sp2.x;
            const double u = h ? u2_.y : u2_.x, um = h ? um2.y : um2.x;
            // u^ma = u^mq u^(2 grp)
            const double uu = u * u;
            const double f = ((grp & 1) ? uu : 1.0) * ((grp & 2) ? uu * uu : 1.0);
            const double pwa = live_a ? um * f : 0.0;
            const double pwb = live_b ? pwa * u : 0.0;
            // gen_point_pressure.py:186-188: exp(1j * (m * phi)) of the ROUNDED product, for both orders.
            // Order ma directly; order ma + 1 by one rotation: fl((ma+1) phi) = fl(ma phi) + phi + d with d
            // formed exactly (both differences are exact), cos(a + phi + d) = cos(a + phi) - d sin(a + phi).
            const double a = dma * ph, a2 = dmb * ph;
            double sv, cv, c1v, s1v;
            if (SAFE) {
              sincos(a, &sv, &cv);
              sincos(a2, &s1v, &c1v);
              c1v *= pwb;
              s1v *= pwb;
            } else {
              pf_sincos(a, sv, cv);
              const double d = (a2 - a) - ph;
              const double c2 = fma(cv, cp, -(sv * sp)), s2 = fma(sv, cp, cv * sp);
              c1v = pwb * fma(-d, s2, c2);                                         // order ma + 1
              s1v = pwb * fma(d, c2, s2);
            }
            const double c0v = pwa * cv, s0v = pwa * sv;                           // order ma
            const double cA = b2 ? c1v : c0v, sA = b2 ? s1v : s0v, cB = b2 ? c0v : c1v, sB = b2 ? s0v : s1v;
            fa[i] = b1 ? sA : cA;
            ga[i] = b1 ? cA : sA;
            fb[i] = b1 ? sB : cB;
            gb[i] = b1 ? cB : sB;
            pa1[i] = b2 ? sgn : 0.0;
            pa2[i] = b2 ? 0.0 : sgn;
            pb1[i] = b2 ? 0.0 : sgn;
            pb2[i] = b2 ? sgn : 0.0;
          }
        }
      }
      const double* wrow = tile + (size_t)mq * TILE + 2 * j;
      // multipliers of this lane's two orders, one row of the table per degree (rows past LMAX and the
      // entries with l <= m are zero, l = m + 1 is one): L1-resident, fetched one iteration ahead
      const double* gpa = q.gl + (size_t)mq * q.Mpad + ma + (b2 ? 1 : 0);
      const double* gpb = q.gl + (size_t)mq * q.Mpad + ma + (b2 ? 0 : 1);
      const size_t gstride = (size_t)q.Mpad;
      double ga0 = __ldg(gpa), gb0 = __ldg(gpb), ga1 = __ldg(gpa + gstride), gb1 = __ldg(gpb + gstride);
      double* ap = acc_w + (size_t)row0 * 16 + acc_lane;
      // one degree: v = {A first, A second, B first, B second}
      auto step = [&](const double gA, const double gB, double (&v)[4]) {
        double w[P];
        pf_load(wrow, w);
        wrow += TILE;
        double a1 = 0.0, a2 = 0.0, b1s = 0.0, b2s = 0.0;
#pragma unroll
        for (int i = 0; i < P; ++i) {
          const double qa = fma(x[i] * gA, pa1[i], -pa2[i]);
          const double qb = fma(x[i] * gB, pb1[i], -pb2[i]);
          pa2[i] = pa1[i];
          pa1[i] = qa;
          pb2[i] = pb1[i];
          pb1[i] = qb;
          const double ta = qa * w[i], tb2 = qb * w[i];
          a1 = fma(ta, fa[i], a1);
          a2 = fma(ta, ga[i], a2);
          b1s = fma(tb2, fb[i], b1s);
          b2s = fma(tb2, gb[i], b2s);
        }
        v[0] = a1; v[1] = a2; v[2] = b1s; v[3] = b2s;
      };
      // Transposed reduction over the duo's 8 lanes: keep slot A, receive the partner's slot B (4 exchanges), keep
      // the first component, receive the partner's second (2), then the step (1).  It runs ONE ITERATION LATE,
      // its exchanges placed between the two steps of the next iteration, so that the shuffle latencies overlap
      // that arithmetic (the first pass adds zeros to a spare row).
      double pa_[4] = {0.0, 0.0, 0.0, 0.0}, pb_[4] = {0.0, 0.0, 0.0, 0.0};
      double* app = acc_w + (size_t)(NS - 2) * 16 + acc_lane;     // spare rows
      for (int s = 0; s < n2; s += 2) {
        gpa += 2 * gstride;
        gpb += 2 * gstride;
        const double ha0 = __ldg(gpa), hb0 = __ldg(gpb);                 // rows up to LMAX + 3 exist
        const double ha1 = __ldg(gpa + gstride), hb1 = __ldg(gpb + gstride);
        double va[4], vb[4];
        const double r1a = __shfl_xor_sync(0xffffffffu, pa_[2], 4), r2a = __shfl_xor_sync(0xffffffffu, pa_[3], 4);
        const double r1b = __shfl_xor_sync(0xffffffffu, pb_[2], 4), r2b = __shfl_xor_sync(0xffffffffu, pb_[3], 4);
        step(ga0, gb0, va);
        const double k1a = pa_[0] + r1a, k2a = pa_[1] + r2a, k1b = pb_[0] + r1b, k2b = pb_[1] + r2b;
        const double r3a = __shfl_xor_sync(0xffffffffu, k2a, 2), r3b = __shfl_xor_sync(0xffffffffu, k2b, 2);
        step(ga1, gb1, vb);
        const double ma_ = k1a + r3a, mb_ = k1b + r3b;
        const double k1 = (b0 ? mb_ : ma_) + __shfl_xor_sync(0xffffffffu, b0 ? ma_ : mb_, 1);
        *app += k1;
        app = ap;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          pa_[i] = va[i];
          pb_[i] = vb[i];
        }
        ga0 = ha0; gb0 = hb0; ga1 = ha1; gb1 = hb1;
        ap += 32;
      }
      {
        const double k1a = pa_[0] + __shfl_xor_sync(0xffffffffu, pa_[2], 4);
        const double k2a = pa_[1] + __shfl_xor_sync(0xffffffffu, pa_[3], 4);
        const double k1b = pb_[0] + __shfl_xor_sync(0xffffffffu, pb_[2], 4);
        const double k2b = pb_[1] + __shfl_xor_sync(0xffffffffu, pb_[3], 4);
        const double ma_ = k1a + __shfl_xor_sync(0xffffffffu, k2a, 2);
        const double mb_ = k1b + __shfl_xor_sync(0xffffffffu, k2b, 2);
        const double k1 = (b0 ? mb_ : ma_) + __shfl_xor_sync(0xffffffffu, b0 ? ma_ : mb_, 1);
        *app += k1;
      }
      row0 += n2;
    }
    pf_bar_arrive(PF_BAR_EMPTY + cur, 256);            // this warp is done with the buffer
  }
  __syncwarp();
  // this CTA's partial sums: the (l >= m) entries of the two walks
  double* slot = q.slots + ((size_t)blockIdx.x * q.nbatch + t) * 2 * (L + 1) * (q.mmax + 1);
  const size_t plane = (size_t)(L + 1) * (q.mmax + 1);
  int row0 = 0;
  for (int wk = 0; wk < 2; ++wk) {
    const int mq = mq2[wk];
    const int n = (mq <= q.mmax) ? (L + 1 - mq) : 0;
    for (int i = lane; i < n * 16; i += 32) {
      const int s = i >> 4, e = i & 15;
      const int l = mq + s, m = mq + (e >> 1);
      if (m <= q.mmax && l >= m)
        slot[(size_t)(e & 1) * plane + (size_t)l * (q.mmax + 1) + m] = acc_w[(size_t)(row0 + s) * 16 + e];
    }
    row0 += n2_[wk];
  }
}

// grid (lmax+1, nbatch, 2), 1024 threads: out = dfactor[l] * sum over CTA partials (z: cosine / sine sums).  Sixteen interleaved
// partial sums per (l, m) keep independent loads in flight; they are combined in a fixed order.
constexpr int PR_GROUPS = 16;
__global__ void __launch_bounds__(64 * PR_GROUPS) points_reduce_kernel(const PointParams q,
                                                                       const __grid_constant__ DegreeFactors dfs) {
  __shared__ double part[PR_GROUPS][64];
  const int l = blockIdx.x, t = blockIdx.y, comp = blockIdx.z;      // comp 0: cosine sums, 1: sine sums
  const int ncol = q.mmax + 1;
  const size_t plane = (size_t)(q.lmax + 1) * ncol;
  const double df = (dfs.dev != nullptr) ? dfs.dev[l] : dfs.v[l];
  const int ml = threadIdx.x & 63, grp = threadIdx.x >> 6;
  double* out = comp ? q.slm : q.clm;
  for (int m0 = 0; m0 < ncol; m0 += 64) {
    const int m = m0 + ml;
    double c = 0.0;
    if (m <= l && m < ncol) {
      const double* src = q.slots + (size_t)t * 2 * plane + (size_t)comp * plane + (size_t)l * ncol + m;
      const size_t kstride = (size_t)q.nbatch * 2 * plane;
#pragma unroll 8
      for (int k = grp; k < q.ncta_p; k += PR_GROUPS) c += src[(size_t)k * kstride];
    }
    part[grp][ml] = c;
    __syncthreads();
    if (grp == 0 && m < ncol) {
      // fixed pairwise tree over the groups
      double v[PR_GROUPS];
#pragma unroll
      for (int j = 0; j < PR_GROUPS; ++j) v[j] = part[j][ml];
#pragma unroll
      for (int w = 1; w < PR_GROUPS; w *= 2) {
#pragma unroll
        for (int j = 0; j < PR_GROUPS; j += 2 * w) v[j] += v[j + w];
      }
      double r = 0.0;
      if (m <= l) {
        r = v[0] * df;
        if (q.scale != nullptr) r *= __ldg(q.scale + (size_t)l * q.Mpad + m);
      }
      out[((size_t)t * (q.lmax + 1) + l) * ncol + m] = r;
    }
    __syncthreads();
  }
}

}  // namespace mhb

using namespace mhb;

namespace {

struct PointLayout {
  long long npad;
  int Mpad, ncta_p, ncta_max, batches_per_cta, n_mtiles;
  size_t off_W, off_x, off_u, off_E, off_df, off_slots, off_flag, total;
};

PointLayout point_layout(long long npts, int nbatch, int lmax, int mmax, int num_sms) {
  PointLayout y;
  y.npad = (npts + PT_BATCH - 1) / PT_BATCH * PT_BATCH;
  y.Mpad = (mmax + 64) / 64 * 64;
  y.n_mtiles = (mmax + PT_MTILE) / PT_MTILE;
  const long long nbatches = y.npad / PT_BATCH;
  const long long want = std::max<long long>(1, 2LL * num_sms / (y.n_mtiles * (long long)nbatch));
  y.ncta_p = (int)std::min<long long>(nbatches, std::max<long long>(want, 1));
  y.batches_per_cta = (int)((nbatches + y.ncta_p - 1) / y.ncta_p);
  y.ncta_p = (int)((nbatches + y.batches_per_cta - 1) / y.batches_per_cta);
  // the fused kernel runs up to three CTAs per SM on tiles of at least 48 points
  y.ncta_max = (int)std::max<long long>(
      y.ncta_p, std::min<long long>((npts + 47) / 48,
                                    std::max<long long>(1, 3LL * num_sms / (y.n_mtiles * (long long)nbatch))));
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += (n + 31) / 32 * 32; return r; };
  y.off_W = take((size_t)nbatch * (lmax + 1) * y.npad);
  y.off_x = take((size_t)y.npad);
  y.off_u = take((size_t)y.npad);
  y.off_E = take(2 * (size_t)(mmax + 1) * y.npad);
  y.off_df = take((size_t)lmax + 1);
  y.off_slots = take((size_t)y.ncta_max * nbatch * 2 * (lmax + 1) * (mmax + 1));
  y.off_flag = take(4);
  y.total = o * sizeof(double);
  return y;
}

int g_sms_cache[64] = {0};

struct PointTabKey {
  int device, lmax, mmax;
  bool operator<(const PointTabKey& o) const {
    return std::tie(device, lmax, mmax) < std::tie(o.device, o.lmax, o.mmax);
  }
};
std::map<PointTabKey, double*> g_point_tabs;
std::mutex g_point_tabs_mu;

// device table [f1 | f2 | pmm | dc1 | dc2], built once
int point_tables(int device, int lmax, int mmax, int Mpad, const double** out) {
  std::lock_guard<std::mutex> lock(g_point_tabs_mu);
  const PointTabKey key{device, lmax, mmax};
  auto it = g_point_tabs.find(key);
  if (it != g_point_tabs.end()) {
    *out = it->second;
    return MHB200_OK;
  }
  const size_t nf = (size_t)(lmax + 1) * Mpad;
  const size_t ng = (size_t)(lmax + 4) * Mpad;          // the fused kernel reads up to three zero rows past LMAX
  std::vector<double> tab(3 * nf + ng + (size_t)Mpad + 2 * (size_t)(lmax + 1), 0.0);
  double *f1 = tab.data(), *f2 = f1 + nf, *pmm = f2 + nf, *dc1 = pmm + Mpad, *dc2 = dc1 + (lmax + 1);
  double *gl = dc2 + (lmax + 1), *sc = gl + ng;
  for (int l = 2; l <= lmax; ++l) {
    const double dl = (double)l;
    f1[(size_t)l * Mpad] = sqrt(2.0 * dl - 1.0) * sqrt(2.0 * dl + 1.0) / dl;
    f2[(size_t)l * Mpad] = (dl - 1.0) * sqrt(2.0 * dl + 1.0) / (sqrt(2.0 * dl - 3.0) * dl);
    for (int m = 1; m <= std::min(l - 2, mmax); ++m) {
      const double dm = (double)m;
      f1[(size_t)l * Mpad + m] = sqrt(2.0 * dl + 1.0) * sqrt(2.0 * dl - 1.0) / (sqrt(dl + dm) * sqrt(dl - dm));
      f2[(size_t)l * Mpad + m] = sqrt(2.0 * dl + 1.0) * sqrt(dl - dm - 1.0) * sqrt(dl + dm - 1.0) /
                                 (sqrt(2.0 * dl - 3.0) * sqrt(dl + dm) * sqrt(dl - dm));
    }
  }
  // seed rows (see points_contract_kernel): degree m reproduces pmm, degree m+1 gives (x*sqrt(2m+3))*pmm
  for (int m = 0; m <= mmax; ++m) {
    f1[(size_t)m * Mpad + m] = 0.0;
    f2[(size_t)m * Mpad + m] = -1.0;
    if (m + 1 <= lmax) {
      f1[(size_t)(m + 1) * Mpad + m] = sqrt(2.0 * m + 3.0);
      f2[(size_t)(m + 1) * Mpad + m] = 0.0;
    }
  }
  pmm[0] = 1.0;
  double v = sqrt(2.0) * 1.0e-280;
  for (int m = 1; m <= mmax; ++m) {
    v = v * sqrt(2.0 * m + 1.0) / sqrt(2.0 * m);
    pmm[m] = v;
  }
  // gen_point_pressure.py:139-141: the quotients of the disc recursion, formed as the reference forms them
  for (int l = 0; l <= lmax; ++l) {
    const double dl = (double)l;
    dc1[l] = (2.0 * dl + 1.0) / (dl + 1.0);
    dc2[l] = dl / (dl + 1.0);
  }
  // one-multiplier form of the same recursion (points_fused_kernel): q_l = Plm / (u^m c_lm) with
  // c_mm = Pmm / u^m, c_(m+1)m = sqrt(2m+3) c_mm, c_lm = f2_lm c_(l-2)m, so that
  // q_m = 1, q_(m+1) = x, q_l = (x g_lm) q_(l-1) - q_(l-2), g_lm = f1_lm c_(l-1)m / c_lm.
  // Formed in extended precision and rounded once.
  {
    long double cmm = 1.0L;
    for (int m = 0; m <= mmax; ++m) {
      if (m == 1) cmm = sqrtl(3.0L);
      if (m > 1) cmm = cmm * sqrtl(2.0L * m + 1.0L) / sqrtl(2.0L * m);
      long double c2 = cmm, c1 = 0.0L;
      sc[(size_t)m * Mpad + m] = (double)cmm;
      if (m + 1 <= lmax) {
        c1 = cmm * sqrtl(2.0L * m + 3.0L);
        sc[(size_t)(m + 1) * Mpad + m] = (double)c1;
        gl[(size_t)(m + 1) * Mpad + m] = 1.0;
      }
      for (int l = m + 2; l <= lmax; ++l) {
        const long double dl = l, dm = m;
        const long double a = sqrtl((2.0L * dl + 1.0L) * (2.0L * dl - 1.0L) / ((dl + dm) * (dl - dm)));
        const long double b = sqrtl((2.0L * dl + 1.0L) * (dl - dm - 1.0L) * (dl + dm - 1.0L) /
                                    ((2.0L * dl - 3.0L) * (dl + dm) * (dl - dm)));
        const long double c = b * c2;
        gl[(size_t)l * Mpad + m] = (double)(a * c1 / c);
        sc[(size_t)l * Mpad + m] = (double)c;
        c2 = c1;
        c1 = c;
      }
    }
  }
  double* dev = nullptr;
  MHB_CUDA(cudaMalloc((void**)&dev, tab.size() * sizeof(double)));
  MHB_CUDA(cudaMemcpy(dev, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
  g_point_tabs[key] = dev;
  *out = dev;
  return MHB200_OK;
}

int sm_count(int device, int* out) {
  if (device >= 0 && device < 64 && g_sms_cache[device]) {
    *out = g_sms_cache[device];
    return MHB200_OK;
  }
  cudaDeviceProp prop;
  MHB_CUDA(cudaGetDeviceProperties(&prop, device));
  *out = prop.multiProcessorCount;
  if (device >= 0 && device < 64) g_sms_cache[device] = *out;
  return MHB200_OK;
}

}  // namespace

extern "C" {

size_t mhb200_points_workspace_bytes(int64_t npts, int nbatch, int lmax, int mmax) {
  if (npts < 1 || nbatch < 1 || lmax < 0 || mmax < 0 || mmax > lmax) return 0;
  // sized for the largest SM count we could meet so the figure does not depend on the device
  return point_layout(npts, nbatch, lmax, mmax, 160).total + 4096;
}

int mhb200_point_pressure_f64(int device, const double* P, const double* G, int64_t G_stride, const double* R,
                              int64_t R_stride, const double* phi, const double* theta, const double* area,
                              int64_t npts, int nbatch, int lmax, int mmax, double rad_e, const double* dfactor,
                              double* clm_out, double* slm_out, void* workspace, size_t workspace_bytes,
                              void* stream) {
  if (!P || !G || !R || !phi || !theta || !area || npts < 1 || nbatch < 1 || lmax < 0 || mmax < 0 ||
      mmax > lmax || !dfactor || !clm_out || !slm_out || !workspace) {
    set_err("mhb200_point_pressure_f64: bad argument");
    return MHB200_EINVAL;
  }
  int rc = mhb200_device_check(device);
  if (rc != MHB200_OK) return rc;
  MHB_CUDA(cudaSetDevice(device));
  int sms = 0;
  if ((rc = sm_count(device, &sms)) != MHB200_OK) return rc;
  if (workspace_bytes < mhb200_points_workspace_bytes(npts, nbatch, lmax, mmax)) {
    set_err("mhb200_point_pressure_f64: workspace too small");
    return MHB200_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const PointLayout y = point_layout(npts, nbatch, lmax, mmax, std::min(sms, 160));
  double* ws = (double*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);

  // Holmes-Featherstone multipliers (same expressions as the grid plan), cached per (device, lmax, mmax)
  const int Mpad = y.Mpad;
  const double* tab = nullptr;
  if ((rc = point_tables(device, lmax, mmax, Mpad, &tab)) != MHB200_OK) return rc;
  DegreeFactors dfs;
  dfs.dev = nullptr;
  if (lmax + 1 <= DF_MAX) {
    for (int l = 0; l <= lmax; ++l) dfs.v[l] = dfactor[l];
  } else {
    // pageable source: the runtime stages it before returning
    MHB_CUDA(cudaMemcpyAsync(ws + y.off_df, dfactor, (size_t)(lmax + 1) * 8, cudaMemcpyHostToDevice, st));
    dfs.dev = ws + y.off_df;
  }

  PointParams q;
  q.npts = npts;
  q.npad = y.npad;
  q.nbatch = nbatch;
  q.lmax = lmax;
  q.mmax = mmax;
  q.P = P;
  q.G = G;
  q.R = R;
  q.phi = phi;
  q.theta = theta;
  q.area = area;
  q.Gs = G_stride;
  q.Rs = R_stride;
  q.rad_e = rad_e;
  q.area_denom = 2.0 * 3.141592653589793 * (rad_e * rad_e);
  q.W = ws + y.off_W;
  q.xg = ws + y.off_x;
  q.ug = ws + y.off_u;
  q.E = reinterpret_cast<double2*>(ws + y.off_E);
  q.f1 = tab;
  q.f2 = tab + (size_t)(lmax + 1) * Mpad;
  q.pmm = q.f2 + (size_t)(lmax + 1) * Mpad;
  q.dc1 = q.pmm + Mpad;
  q.dc2 = q.dc1 + (lmax + 1);
  q.Mpad = Mpad;
  q.batches_per_cta = y.batches_per_cta;
  q.slots = ws + y.off_slots;
  q.ncta_p = y.ncta_p;
  q.clm = clm_out;
  q.slm = slm_out;
  q.gl = q.dc2 + (lmax + 1);
  q.scale = nullptr;
  q.tiles_per_cta = 0;
  q.tiles_rem = 0;
  q.fused_steps = 0;
  q.flag = nullptr;
  q.gen = 0;

  // fused kernel (one launch from the point arrays to the partial sums) wherever two of its CTAs fit on an SM
  {
    int steps = 0;
    for (int j = 0; j < y.n_mtiles; ++j)
      for (int w = 0; w < 4; ++w)
        steps = std::max(steps, fused_walk_steps(lmax, mmax, j * PF_ORD, w) +
                                    fused_walk_steps(lmax, mmax, j * PF_ORD, 7 - w));
    steps += 2;                                                // two spare accumulator rows
    const char* env = getenv("MHB200_POINTS_FUSED");
    const bool want = !(env && env[0] == '0');
    const size_t smem_fused = ((size_t)4 * steps * 16 + 2 * ((lmax + 2) & ~1) +
                               2 * (size_t)(lmax + 3 + PF_AUX) * PF_TILE) * sizeof(double);
    if (want && smem_fused <= (size_t)(228 * 1024) / 2 - 1024) {
      const long long ntiles = (npts + PF_TILE - 1) / PF_TILE;
      const long long cap = std::max<long long>(1, 2LL * sms / (y.n_mtiles * (long long)nbatch));
      const long long ctas = std::max<long long>(1, std::min<long long>(ntiles, std::min<long long>(cap, y.ncta_max)));
      q.tiles_per_cta = (int)(ntiles / ctas);
      q.tiles_rem = (int)(ntiles % ctas);
      q.ncta_p = (int)ctas;
      q.fused_steps = steps;
      q.scale = q.gl + (size_t)(lmax + 4) * Mpad;
      dim3 grid(q.ncta_p, y.n_mtiles, nbatch);
      profile_begin(st);
      static std::atomic<unsigned long long> generation{0x9e3779b97f4a7c15ull};
      q.flag = reinterpret_cast<unsigned long long*>(ws + y.off_flag);
      q.gen = generation.fetch_add(0x9e3779b97f4a7c15ull);
      MHB_CUDA(cudaFuncSetAttribute(points_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem_fused));
      MHB_CUDA(cudaFuncSetAttribute(points_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem_fused));
      points_fused_kernel<false><<<grid, PF_NT, smem_fused, st>>>(q);
      profile_end(st);
      points_fused_kernel<true><<<grid, PF_NT, smem_fused, st>>>(q);     // returns at once unless flagged
      g_launches.fetch_add(2);
      MHB_CUDA(cudaGetLastError());
      dim3 rgrid(lmax + 1, nbatch, 2);
      points_reduce_kernel<<<rgrid, 64 * PR_GROUPS, 0, st>>>(q, dfs);
      g_launches.fetch_add(1);
      MHB_CUDA(cudaGetLastError());
      return MHB200_OK;
    }
  }

  {
    const int nmg = (mmax + PT_MGROUP) / PT_MGROUP;
    dim3 grid((unsigned)((y.npad + 127) / 128), nmg + nbatch);
    points_prepare_kernel<<<grid, 128, 0, st>>>(q, nmg);
    g_launches.fetch_add(1);
    MHB_CUDA(cudaGetLastError());
  }
  {
    int pitch_max = 0;
    for (int j = 0; j < y.n_mtiles; ++j) {
      const int a = j * PT_MTILE, b = std::min(a + PT_MTILE - 1, mmax);
      pitch_max = std::max(pitch_max, 2 * (lmax + 1) - (a + b) + 1);
    }
    const size_t smem_acc = (size_t)PT_PAIRS * 2 * pitch_max * sizeof(double);
    const int fpitch_max = staged_fpitch(pitch_max - 1);
    const size_t smem_staged = smem_acc + (size_t)PT_PAIRS * fpitch_max * 16 +
                               (size_t)2 * (lmax + 2) * WT_PITCH * sizeof(double) + (size_t)PT_NT * PT_PER_LANE * 16;
    const size_t smem_max = 227 * 1024;
    if (smem_acc > smem_max) {
      set_err("mhb200_point_pressure_f64: lmax=%d needs %zu bytes of shared memory", lmax, smem_acc);
      return MHB200_EINVAL;
    }
    dim3 grid(y.ncta_p, y.n_mtiles, nbatch);
    profile_begin(st);
    // the staged kernel only where two of its CTAs fit on an SM (each CTA also reserves 1 KB): with
    // one CTA (4 warps) per SM the direct-load kernel at two CTAs hides its latencies better
    const size_t smem_two_ctas = (228 * 1024) / 2 - 1024;
    if (smem_staged <= smem_two_ctas) {
      MHB_CUDA(cudaFuncSetAttribute(points_contract_staged, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem_staged));
      points_contract_staged<<<grid, PT_NT, smem_staged, st>>>(q);
    } else {
      MHB_CUDA(cudaFuncSetAttribute(points_contract_generic, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem_acc));
      points_contract_generic<<<grid, PT_NT, smem_acc, st>>>(q);
    }
    profile_end(st);
    g_launches.fetch_add(1);
    MHB_CUDA(cudaGetLastError());
  }
  {
    dim3 grid(lmax + 1, nbatch, 2);
    points_reduce_kernel<<<grid, 64 * PR_GROUPS, 0, st>>>(q, dfs);
    g_launches.fetch_add(1);
    MHB_CUDA(cudaGetLastError());
  }
  return MHB200_OK;
}

// test hook: the fused point kernel's short sincos on n device values (|a| < 5e5)
int mhb200_sincos_probe(int device, const double* a, int64_t n, double* sin_out, double* cos_out, void* stream) {
  if (!a || !sin_out || !cos_out || n < 1) {
    set_err("mhb200_sincos_probe: bad argument");
    return MHB200_EINVAL;
  }
  int rc = mhb200_device_check(device);
  if (rc != MHB200_OK) return rc;
  MHB_CUDA(cudaSetDevice(device));
  sincos_probe_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, n, sin_out, cos_out);
  MHB_CUDA(cudaGetLastError());
  return MHB200_OK;
}

}  // extern "C"
