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
// Two kernels:
//  (1) points_weights_kernel: one thread per point walks the degrees.  The disc factor
//      Pdisc_l = (P_{l-1}(alpha) - P_{l+1}(alpha))/2 is a cancellation-prone 3-term recursion
//      (error amplified by ~1/(1-alpha) ~ npts/2), so it is evaluated in EXACTLY the
//      reference's operation order with round-to-nearest intrinsics and no FMA contraction.
//  (2) points_contract_kernel: scatter-reduce.  A lane group of 4 lanes owns one pair of
//      orders (m, mtop-m) -- pairing makes every lane walk the same number of degrees -- and
//      each lane carries 8 points at a time through the Holmes-Featherstone recursion in
//      registers (Plm normalisation identical to legendre(NORMALIZE=True)); the 4 lanes are
//      combined with warp shuffles and added to accumulators private to the lane group in
//      shared memory.  Per-CTA partials go to slots that a final kernel reduces in fixed order
//      (deterministic; no fp64 atomics).
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
  double2* E;                   // [mmax+1][npad] (u^m/1e-280) * (cos m phi, sin m phi)
  // contraction
  const double *f1, *f2, *pmm, *sq;   // Holmes-Featherstone multipliers [(lmax+1)][Mpad]
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

// gen_point_pressure.py:126-147
__global__ void points_weights_kernel(const PointParams q) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int t = blockIdx.y;
  if (p >= q.npad) return;
  double* W = q.W + (size_t)t * (q.lmax + 1) * q.npad + p;
  if (p >= q.npts) {
    for (int l = 0; l <= q.lmax; ++l) W[(size_t)l * q.npad] = 0.0;
    if (t == 0) q.xg[p] = 0.0;
    return;
  }
  if (t == 0) q.xg[p] = cos(q.theta[p]);
  const double alpha = __dsub_rn(1.0, __ddiv_rn(q.area[p], q.area_denom));
  const double PG = __ddiv_rn(q.P[(size_t)t * q.npts + p], q.G[p * q.Gs]);
  const double r = __ddiv_rn(q.R[p * q.Rs], q.rad_e);
  double Pm1 = 1.0, Pl = 1.0;
  double pw = __dmul_rn(r, r);
  for (int l = 0; l <= q.lmax; ++l) {
    const double dl = (double)l;
    const double c1 = __ddiv_rn(__dadd_rn(__dmul_rn(2.0, dl), 1.0), __dadd_rn(dl, 1.0));
    const double c2 = __ddiv_rn(dl, __dadd_rn(dl, 1.0));
    const double Pp1 = __dsub_rn(__dmul_rn(__dmul_rn(c1, alpha), Pl), __dmul_rn(c2, Pm1));
    const double Pdisc = __dmul_rn(__dsub_rn(Pm1, Pp1), 0.5);
    W[(size_t)l * q.npad] = __dmul_rn(__dmul_rn(Pdisc, PG), pw);
    pw = __dmul_rn(pw, r);
    Pm1 = Pl;
    Pl = Pp1;
  }
}

// Per (point, order) factors of the contraction, one thread each (throughput-oriented: the
// contraction kernel used to run these sincos / power chains serially in low-occupancy warps):
//   E[m][p] = (u_p^m / 1e-280) * (cos(m phi_p), sin(m phi_p)),  u = sqrt(1 - x^2) (0 -> eps)
// i.e. e^{i m phi} of gen_point_pressure.py:186-188 with the rescale of the scaled
// Holmes-Featherstone column folded in.  grid (npad/128, mmax+1).
__global__ void points_trig_kernel(const PointParams q) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (p >= q.npad) return;
  double2 e = make_double2(0.0, 0.0);
  if (p < q.npts) {
    const double x = cos(q.theta[p]);
    double u = sqrt(1.0 - x * x);
    if (u == 0.0) u = 2.220446049250313e-16;
    double r = (m == 0) ? 1.0 : 1.0e280, bb = u;
    int k = m;
    while (k) {
      if (k & 1) r *= bb;
      bb *= bb;
      k >>= 1;
    }
    double sv, cv;
    sincos((double)m * q.phi[p], &sv, &cv);
    e = make_double2(r * cv, r * sv);
  }
  q.E[(size_t)m * q.npad + p] = e;
}

// grid: (ncta_p, n_mtiles, nbatch)
__global__ void __launch_bounds__(PT_NT, 2) points_contract_kernel(const PointParams q) {
  extern __shared__ double acc_s[];   // [PT_PAIRS][2][pitch]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane & (PT_GROUPS - 1);
  const int pair = warp * (32 / PT_GROUPS) + lane / PT_GROUPS;
  // lane groups of one warp walk different degree ranges: shuffles name only the group's lanes
  const unsigned gmask = ((1u << PT_GROUPS) - 1u) << (lane & ~(PT_GROUPS - 1));
  const int t = blockIdx.z;
  const int L = q.lmax;
  // orders of this CTA's tile and of this lane group's pair
  const int m0 = blockIdx.y * PT_MTILE;
  const int m1 = min(m0 + PT_MTILE - 1, q.mmax);
  const int ma = m0 + pair, mb = m1 - pair;
  const bool has_a = (ma <= mb);          // pair exists
  const bool has_b = (ma < mb);           // second order distinct
  const int pitch = 2 * (L + 1) - (m0 + m1) + 1;
  double* acc_c = acc_s + (size_t)pair * 2 * pitch;
  double* acc_sn = acc_c + pitch;
  for (int i = tid; i < PT_PAIRS * 2 * pitch; i += PT_NT) acc_s[i] = 0.0;
  __syncthreads();

  const double* Wt = q.W + (size_t)t * (L + 1) * q.npad;
  const long long batch0 = (long long)blockIdx.x * q.batches_per_cta;
  const long long nbatches = (q.npad + PT_BATCH - 1) / PT_BATCH;
  const long long batch1 = min(batch0 + (long long)q.batches_per_cta, nbatches);

  if (has_a) {
    for (long long b = batch0; b < batch1; ++b) {
      const long long pbase = b * PT_BATCH + g * PT_PER_LANE;
      double x[PT_PER_LANE];
#pragma unroll
      for (int i = 0; i < PT_PER_LANE; ++i) x[i] = q.xg[pbase + i];      // npad is a multiple of PT_BATCH
      int step = 0;
      for (int which = 0; which < 2; ++which) {
        if (which == 1 && !has_b) break;
        const int m = which ? mb : ma;
        // per (point, order) trig/rescale factors from the precomputed table, and the recursion seeds
        double cm[PT_PER_LANE], sm[PT_PER_LANE], p1[PT_PER_LANE], p2[PT_PER_LANE];
        const double pmm = q.pmm[m], sqm = q.sq[m];
        const double2* Em = q.E + (size_t)m * q.npad + pbase;
#pragma unroll
        for (int i = 0; i < PT_PER_LANE; ++i) {
          const double2 e = __ldg(Em + i);
          cm[i] = e.x;
          sm[i] = e.y;
          p2[i] = pmm;                      // degree m
          p1[i] = (x[i] * sqm) * pmm;       // degree m+1
        }
        // Degree loop.  W rows and recursion multipliers live in two register sets used on
        // alternate degrees (loop unrolled by two, no rotation copies, so a load issued right after
        // a set is used has a full degree of work to land), with L1 prefetches four rows ahead.
        auto emit = [&](const double(&pl)[PT_PER_LANE], const double4& w0, const double4& w1) {
          const double w[PT_PER_LANE] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
          // two interleaved partial sums per component keep the FMA chains short
          double c0 = 0.0, c1 = 0.0, s0 = 0.0, s1 = 0.0;
#pragma unroll
          for (int i = 0; i < PT_PER_LANE; i += 2) {
            const double t0 = pl[i] * w[i], t1 = pl[i + 1] * w[i + 1];
            c0 = fma(t0, cm[i], c0);
            c1 = fma(t1, cm[i + 1], c1);
            s0 = fma(t0, sm[i], s0);
            s1 = fma(t1, sm[i + 1], s1);
          }
          double c = c0 + c1, s2 = s0 + s1;
          // combine the lanes of the group (fixed butterfly order)
#pragma unroll
          for (int off = PT_GROUPS / 2; off >= 1; off >>= 1) {
            c += __shfl_xor_sync(gmask, c, off);
            s2 += __shfl_xor_sync(gmask, s2, off);
          }
          if (g == 0) {
            acc_c[step] += c;
            acc_sn[step] += s2;
          }
          ++step;
        };
        const double* f1 = q.f1 + m;
        const double* f2 = q.f2 + m;
        struct RowSet {
          double4 w0, w1;
          double a, b;
        };
        auto load_row = [&](RowSet& r, int l) {
          const int ln = min(l, L);
          const double* Wn = Wt + (size_t)ln * q.npad + pbase;
          r.w0 = *reinterpret_cast<const double4*>(Wn);
          r.w1 = *reinterpret_cast<const double4*>(Wn + 4);
          r.a = __ldg(f1 + (size_t)ln * q.Mpad);
          r.b = __ldg(f2 + (size_t)ln * q.Mpad);
          const double* Wp = Wt + (size_t)min(l + 4, L) * q.npad + pbase;
          asm volatile("prefetch.global.L1 [%0];" ::"l"(Wp));
        };
        auto recur = [&](double(&pl)[PT_PER_LANE], const RowSet& r) {
#pragma unroll
          for (int i = 0; i < PT_PER_LANE; ++i) {
            pl[i] = fma(x[i] * r.a, p1[i], -(r.b * p2[i]));
            p2[i] = p1[i];
            p1[i] = pl[i];
          }
        };
        RowSet ra, rb;
        load_row(ra, m);
        load_row(rb, m + 1);
        // degree m
        emit(p2, ra.w0, ra.w1);
        load_row(ra, m + 2);
        if (m + 1 <= L) {
          // degree m+1
          emit(p1, rb.w0, rb.w1);
          load_row(rb, m + 3);
          int l = m + 2;
          for (; l + 1 <= L; l += 2) {
            double pl[PT_PER_LANE];
            recur(pl, ra);
            emit(pl, ra.w0, ra.w1);
            load_row(ra, l + 2);
            recur(pl, rb);
            emit(pl, rb.w0, rb.w1);
            load_row(rb, l + 3);
          }
          if (l <= L) {
            double pl[PT_PER_LANE];
            recur(pl, ra);
            emit(pl, ra.w0, ra.w1);
          }
        }
      }
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

// grid (lmax+1, nbatch), 256 threads: out = dfactor[l] * sum over CTA partials.  Four interleaved
// partial sums per (l, m) keep independent loads in flight; they are combined in a fixed order.
__global__ void __launch_bounds__(256) points_reduce_kernel(const PointParams q,
                                                            const __grid_constant__ DegreeFactors dfs) {
  __shared__ double part[2][4][64];
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
      for (int k = grp; k < q.ncta_p; k += 4) {
        const double* slot = q.slots + ((size_t)k * q.nbatch + t) * 2 * plane;
        c += slot[(size_t)l * ncol + m];
        s += slot[plane + (size_t)l * ncol + m];
      }
    }
    part[0][grp][ml] = c;
    part[1][grp][ml] = s;
    __syncthreads();
    if (grp == 0 && m < ncol) {
      const double cc = ((part[0][0][ml] + part[0][1][ml]) + (part[0][2][ml] + part[0][3][ml])) * df;
      const double ss = ((part[1][0][ml] + part[1][1][ml]) + (part[1][2][ml] + part[1][3][ml])) * df;
      const size_t o = ((size_t)t * (q.lmax + 1) + l) * ncol + m;
      q.clm[o] = (m <= l) ? cc : 0.0;
      q.slm[o] = (m <= l) ? ss : 0.0;
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
  size_t off_W, off_x, off_E, off_df, off_slots, total;
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

// device table [f1 | f2 | pmm | sq], built once
int point_tables(int device, int lmax, int mmax, int Mpad, const double** out) {
  std::lock_guard<std::mutex> lock(g_point_tabs_mu);
  const PointTabKey key{device, lmax, mmax};
  auto it = g_point_tabs.find(key);
  if (it != g_point_tabs.end()) {
    *out = it->second;
    return MHB200_OK;
  }
  const size_t nf = (size_t)(lmax + 1) * Mpad;
  std::vector<double> tab(2 * nf + 2 * (size_t)Mpad, 0.0);
  double *f1 = tab.data(), *f2 = f1 + nf, *pmm = f2 + nf, *sq = pmm + Mpad;
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
  pmm[0] = 1.0;
  double v = sqrt(2.0) * 1.0e-280;
  for (int m = 1; m <= mmax; ++m) {
    v = v * sqrt(2.0 * m + 1.0) / sqrt(2.0 * m);
    pmm[m] = v;
  }
  for (int m = 0; m <= mmax; ++m) sq[m] = sqrt(2.0 * m + 3.0);
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
  q.E = reinterpret_cast<double2*>(ws + y.off_E);
  q.f1 = tab;
  q.f2 = tab + (size_t)(lmax + 1) * Mpad;
  q.pmm = q.f2 + (size_t)(lmax + 1) * Mpad;
  q.sq = q.pmm + Mpad;
  q.Mpad = Mpad;
  q.batches_per_cta = y.batches_per_cta;
  q.slots = ws + y.off_slots;
  q.ncta_p = y.ncta_p;
  q.clm = clm_out;
  q.slm = slm_out;

  {
    dim3 grid((unsigned)((y.npad + 127) / 128), nbatch);
    points_weights_kernel<<<grid, 128, 0, st>>>(q);
    g_launches.fetch_add(1);
    MHB_CUDA(cudaGetLastError());
    dim3 tgrid((unsigned)((y.npad + 127) / 128), mmax + 1);
    points_trig_kernel<<<tgrid, 128, 0, st>>>(q);
    g_launches.fetch_add(1);
    MHB_CUDA(cudaGetLastError());
  }
  {
    const int m0max = (y.n_mtiles - 1) * PT_MTILE;
    int pitch_max = 0;
    for (int j = 0; j < y.n_mtiles; ++j) {
      const int a = j * PT_MTILE, b = std::min(a + PT_MTILE - 1, mmax);
      pitch_max = std::max(pitch_max, 2 * (lmax + 1) - (a + b) + 1);
    }
    (void)m0max;
    const size_t smem = (size_t)PT_PAIRS * 2 * pitch_max * sizeof(double);
    if (smem > 227 * 1024) {
      set_err("mhb200_point_pressure_f64: lmax=%d needs %zu bytes of shared memory", lmax, smem);
      return MHB200_EINVAL;
    }
    MHB_CUDA(cudaFuncSetAttribute(points_contract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(y.ncta_p, y.n_mtiles, nbatch);
    profile_begin(st);
    points_contract_kernel<<<grid, PT_NT, smem, st>>>(q);
    profile_end(st);
    g_launches.fetch_add(1);
    MHB_CUDA(cudaGetLastError());
  }
  {
    dim3 grid(lmax + 1, nbatch);
    points_reduce_kernel<<<grid, 256, 0, st>>>(q, dfs);
    g_launches.fetch_add(1);
    MHB_CUDA(cudaGetLastError());
  }
  return MHB200_OK;
}

}  // extern "C"
