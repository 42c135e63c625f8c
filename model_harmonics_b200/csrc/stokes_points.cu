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
// Two kernels + a fixed-order reduce (the two preparation roles share one launch, points_prepare_kernel):
//  (1) weights role: one thread per point walks the degrees.  The disc factor
//      Pdisc_l = (P_{l-1}(alpha) - P_{l+1}(alpha))/2 is a cancellation-prone 3-term recursion
//      (error amplified by ~1/(1-alpha) ~ npts/2), so it is evaluated in EXACTLY the
//      reference's operation order with round-to-nearest intrinsics and no FMA contraction
//      (the degree-only quotients (2l+1)/(l+1), l/(l+1) come from a host table: IEEE division
//      gives the same bits on both sides).
//  (2) trig role: the per (point, order) factors (u^m/1e-280) e^{i m phi}, eight orders per thread.
//  (3) points_contract_kernel: scatter-reduce.  A lane group of 4 lanes owns one pair of
//      orders (m, mtop-m) -- pairing gives every group the same number of degree steps -- and
//      each lane carries 8 points at a time through the Holmes-Featherstone recursion in
//      registers (Plm normalisation identical to legendre(NORMALIZE=True)).  All lanes of a warp
//      run ONE loop over the step index in lock step (a group changes from its first to its
//      second order at its own step, by reloading its state), so the cross-lane sums are
//      full-warp shuffles in uniform control flow.  Sums go to accumulators private to the lane
//      group in shared memory; per-CTA partials go to slots that a final kernel reduces in fixed
//      order (deterministic; no fp64 atomics).
#include "../../include/mhb200.h"
#include "common.cuh"

#include <algorithm>
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

// grid (lmax+1, nbatch), 1024 threads: out = dfactor[l] * sum over CTA partials.  Sixteen interleaved
// partial sums per (l, m) keep independent loads in flight; they are combined in a fixed order.
constexpr int PR_GROUPS = 16;
__global__ void __launch_bounds__(64 * PR_GROUPS) points_reduce_kernel(const PointParams q,
                                                                       const __grid_constant__ DegreeFactors dfs) {
  __shared__ double part[2][PR_GROUPS][64];
  const int l = blockIdx.x, t = blockIdx.y;
  const int ncol = q.mmax + 1;
  const size_t plane = (size_t)(q.lmax + 1) * ncol;
  const double df = (dfs.dev != nullptr) ? dfs.dev[l] : dfs.v[l];
  const int ml = threadIdx.x & 63, grp = threadIdx.x >> 6;
  for (int m0 = 0; m0 < ncol; m0 += 64) {
    const int m = m0 + ml;
    double c = 0.0, s = 0.0;
    if (m <= l && m < ncol) {
#pragma unroll 4
      for (int k = grp; k < q.ncta_p; k += PR_GROUPS) {
        const double* slot = q.slots + ((size_t)k * q.nbatch + t) * 2 * plane;
        c += slot[(size_t)l * ncol + m];
        s += slot[plane + (size_t)l * ncol + m];
      }
    }
    part[0][grp][ml] = c;
    part[1][grp][ml] = s;
    __syncthreads();
    if (grp < 2 && m < ncol) {
      // group 0 finishes the cosine sums, group 1 the sine sums: fixed pairwise tree
      double v[PR_GROUPS];
#pragma unroll
      for (int j = 0; j < PR_GROUPS; ++j) v[j] = part[grp][j][ml];
#pragma unroll
      for (int w = 1; w < PR_GROUPS; w *= 2) {
#pragma unroll
        for (int j = 0; j < PR_GROUPS; j += 2 * w) v[j] += v[j + w];
      }
      const size_t o = ((size_t)t * (q.lmax + 1) + l) * ncol + m;
      const double r = (m <= l) ? v[0] * df : 0.0;
      if (grp == 0) q.clm[o] = r; else q.slm[o] = r;
    }
    __syncthreads();
  }
}

}  // namespace mhb

using namespace mhb;

namespace {

struct PointLayout {
  long long npad;
  int Mpad, ncta_p, batches_per_cta, n_mtiles;
  size_t off_W, off_x, off_u, off_E, off_df, off_slots, total;
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
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += (n + 31) / 32 * 32; return r; };
  y.off_W = take((size_t)nbatch * (lmax + 1) * y.npad);
  y.off_x = take((size_t)y.npad);
  y.off_u = take((size_t)y.npad);
  y.off_E = take(2 * (size_t)(mmax + 1) * y.npad);
  y.off_df = take((size_t)lmax + 1);
  y.off_slots = take((size_t)y.ncta_p * nbatch * 2 * (lmax + 1) * (mmax + 1));
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
  std::vector<double> tab(2 * nf + (size_t)Mpad + 2 * (size_t)(lmax + 1), 0.0);
  double *f1 = tab.data(), *f2 = f1 + nf, *pmm = f2 + nf, *dc1 = pmm + Mpad, *dc2 = dc1 + (lmax + 1);
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
    dim3 grid(lmax + 1, nbatch);
    points_reduce_kernel<<<grid, 64 * PR_GROUPS, 0, st>>>(q, dfs);
    g_launches.fetch_add(1);
    MHB_CUDA(cudaGetLastError());
  }
  return MHB200_OK;
}

}  // extern "C"
