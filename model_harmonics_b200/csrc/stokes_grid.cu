// Gridded fields -> Stokes coefficients on B200 (sm_100a).
//
// Replaces the hot loops of the reference (paths relative to the reference tree):
//   model_harmonics/gen_pressure_stokes.py:194-206      per-degree power pass + complex einsum
//   model_harmonics/gen_atmosphere_stokes.py:217-275    geometry pre-pass, level loop, einsum
// and the third-party gravity_toolkit.plm_holmes table they consume
// (gen_pressure_stokes.py:180-188), which is regenerated on the fly here.
//
// Math, per field:
//   C_lm + i S_lm = D_l * sum_h w_h Plm(x_h) * sum_p e^{i m phi_p} xi_l(h,p)
// One latitude ring h at a time, the inner sum is a dense real contraction
//   Dring[2(M+1) x (L+1)] = Trig[2(M+1) x nlon] * B_h[nlon x (L+1)],  B_h[p,l] = xi_l(h,p)
// done with DMMA m8n8k4 tiles; B_h is *generated* in shared memory (never in HBM):
//   pressure  : (P/G) (R/rad_e)^(l+2), running product in l
//   atmosphere: sum over levels of (dp/gamma)(R/rad_e)^(l+2), geometry fused, one pass over
//               the (level,lat,lon) cubes, coalesced along lon
// then C += w_h Plm(x_h) o Dring with Plm from a per-ring Holmes-Featherstone recursion.
//
// Kernel shape: persistent CTAs (one per SM, 256 threads), static schedule of
// (field, degree-block, order-block, ring-range) segments built on the host so every CTA
// gets equal fp64-pipe work; each segment ends in a private partial-sum slot and a second
// kernel reduces the slots in fixed order (bit-reproducible, no fp64 atomics).
// DFMA and DMMA share one fp64 pipe on B200 (measured: profiles/r01_fp64_peaks.jsonl), so the
// CTA alternates a produce phase (all warps DFMA) and a contraction phase (all warps DMMA)
// rather than specialising warps.
#include "../../include/mhb200.h"
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

namespace mhb {

constexpr int NT = 256;           // threads per CTA
constexpr int LBLK = 64;          // degrees per block
constexpr int MBLK = 64;          // orders per block
constexpr int KC_MAX = 128;       // longitudes per chunk (upper bound)
constexpr int KS_MAX = KC_MAX / 4;
constexpr int PLM_PITCH = 65;
constexpr int CS_ROWS = 36;       // per-thread accumulators kept in shared memory
constexpr int SLOT_DOUBLES = 2 * LBLK * MBLK;
constexpr int DF_MAX = 448;
constexpr int LEV_GROUP = 40;      // levels staged per geometry pass
// region 0 holds the B chunk (64 KB) and, before it is written, the staged (u0, r) pairs of the
// atmosphere geometry pass (LEV_GROUP x 128 x 16 B = 80 KB)
constexpr int REGION0_DOUBLES = LEV_GROUP * KC_MAX * 2;
static_assert(REGION0_DOUBLES >= KS_MAX * 256, "region 0 must hold the B chunk");
constexpr size_t SMEM_BYTES = (size_t)(REGION0_DOUBLES + CS_ROWS * NT + MBLK * PLM_PITCH) * sizeof(double);

enum Mode { MODE_PRESSURE = 0, MODE_ATM_BC05 = 1, MODE_ATM_SW02 = 2 };

// work unit = (ring, longitude chunk), linearised as u = ring * nchunks + chunk
struct Segment {
  int t, lb, mb, u0, u1, pad;
};

struct GridParams {
  int nlat, nlon, lmax, mmax, Mpad, kc, nchunks, nks_total;
  const double *trig, *f1, *f2, *pmm, *sq, *rescale, *x, *w;
  const double* plm_table;
  const Segment* segs;
  const int* cta_seg_begin;
  double* slots;
  // pressure
  const double *P, *G, *R;
  long long Ps[3], Gs[3], Rs[3];
  double rad_e;
  // atmosphere
  const void *GPH, *dP;
  long long batch_stride;
  const double* field_scale;
  const double* geoid;
  long long geo_s0, geo_s1;
  const double *ring_tab, *lon_tab;
  int nlev;
};

struct ReduceParams {
  int lmax, mmax, nmb, ntiles;
  const int* tile_slot_begin;   // [nbatch*ntiles + 1]
  const int* lb_first_tile;     // [nlb]
  const double* slots;
  const double* dfactor_dev;    // used when lmax+1 > DF_MAX
  double *clm, *slm;
  double dfactor[DF_MAX];       // degree factors travel as kernel parameters
};

// ---------------------------------------------------------------------------------------
// Holmes-Featherstone column recursion for one ring and one (degree-block, order-block) tile.
// Thread m_local computes order m = 64*mb + m_local for degrees m..lend and stores the
// weighted values w_h * Plm(x_h) for degrees in the block.  Operation order mirrors
// gravity_toolkit.plm_holmes (oracle/gravtk_standin.py): p_l = (x*f1)*p_{l-1} - f2*p_{l-2},
// unfused, then *rescale (= u^m / 1e-280), then *int_fact.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void legendre_ring(const GridParams& p, int h, int lb, int mb, int m_local,
                                              double* plm_s) {
  const int m = mb * MBLK + m_local;
  const int lbase = lb * LBLK;
  const int lend = min(lbase + LBLK - 1, p.lmax);
  double* row = plm_s + m_local * PLM_PITCH;
  if (p.plm_table != nullptr) {
    const double wh = p.w[h];
    const double* tab = p.plm_table + (size_t)h * (p.lmax + 1) * (p.mmax + 1);
    for (int j = 0; j < LBLK; ++j) {
      const int l = lbase + j;
      double v = 0.0;
      if (l <= p.lmax && m <= p.mmax && m <= l) v = __dmul_rn(tab[(size_t)l * (p.mmax + 1) + m], wh);
      row[j] = v;
    }
    return;
  }
  if (m > p.mmax || m > lend) {
    for (int j = 0; j < LBLK; ++j) row[j] = 0.0;
    return;
  }
  const double x = p.x[h];
  const double wh = p.w[h];
  const double rs = p.rescale[(size_t)h * p.Mpad + m];
  // zero the part of the block below the diagonal and above lmax
  for (int j = 0; j < LBLK; ++j) {
    const int l = lbase + j;
    if (l < m || l > p.lmax) row[j] = 0.0;
  }
  double p2 = p.pmm[m];  // degree m (sectoral), scaled
  if (m >= lbase) row[m - lbase] = __dmul_rn(__dmul_rn(p2, rs), wh);
  if (m + 1 > lend) return;
  double p1 = __dmul_rn(__dmul_rn(x, p.sq[m]), p2);  // degree m+1
  if (m + 1 >= lbase) row[m + 1 - lbase] = __dmul_rn(__dmul_rn(p1, rs), wh);
  const double* f1 = p.f1 + m;
  const double* f2 = p.f2 + m;
  for (int l0 = m + 2; l0 <= lend; l0 += 8) {
    // multipliers for 8 degrees are fetched together so their L2 latency overlaps
    double a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int l = min(l0 + i, lend);
      a[i] = __ldg(f1 + (size_t)l * p.Mpad);
      b[i] = __ldg(f2 + (size_t)l * p.Mpad);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int l = l0 + i;
      if (l <= lend) {
        const double pl = __dsub_rn(__dmul_rn(__dmul_rn(x, a[i]), p1), __dmul_rn(b[i], p2));
        p2 = p1;
        p1 = pl;
        if (l >= lbase) row[l - lbase] = __dmul_rn(__dmul_rn(pl, rs), wh);
      }
    }
  }
}

// B-fragment slot of element (l_local, column) inside the chunk buffer.  Fragment-major:
// [kstep][degree-tile][32 slots]; the row is rotated by kstep so that both the producer's
// stores (16 columns x fixed degree per half-warp) and the consumer's fragment loads are
// bank-conflict free.
__device__ __forceinline__ int bs_index(int l_local, int col) {
  const int ks = col >> 2, c = col & 3;
  return ((ks * 8 + (l_local >> 3)) << 5) + ((((l_local & 7) + ks) & 7) << 2) + c;
}

// ---------------------------------------------------------------------------------------
// Produce phase, pressure:  B[p,l] = (P/G) * (R/rad_e)^(l+2)   (gen_pressure_stokes.py:200)
// A column is shared by the lane pair (lane, lane^16): each handles 32 of the 64 degrees.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void produce_pressure(const GridParams& p, int t, int h, int lb, int chunk,
                                                 double* Bs) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = warp * 16 + (lane & 15), half = lane >> 4;
  if (col >= p.kc) return;
  const int pcol = chunk * p.kc + col;
  double f = 0.0, r = 1.0;
  if (pcol < p.nlon) {
    const double Pv = __ldg(p.P + t * p.Ps[0] + h * p.Ps[1] + pcol * p.Ps[2]);
    const double Gv = __ldg(p.G + t * p.Gs[0] + h * p.Gs[1] + pcol * p.Gs[2]);
    const double Rv = __ldg(p.R + t * p.Rs[0] + h * p.Rs[1] + pcol * p.Rs[2]);
    f = Pv / Gv;
    r = Rv / p.rad_e;
  }
  const int l0 = 32 * half;
  double val = f * ipow(r, lb * LBLK + l0 + 2);
#pragma unroll 8
  for (int j = 0; j < 32; ++j) {
    Bs[bs_index(l0 + j, col)] = val;
    val *= r;
  }
}

// Branch-free fp64 reciprocal and square root (Newton from the fp32 SFU seed).  The library
// sqrt()/division carry special-case branches that split the basic block and serialise the
// geometry chain; operands here are always normal and far from the range limits.
__device__ __forceinline__ double fast_rcp(double a) {
  double y = (double)__frcp_rn((float)a);
  double e = fma(-a, y, 1.0);
  y = fma(y, e, y);
  e = fma(-a, y, 1.0);
  return fma(y, e, y);
}
__device__ __forceinline__ double fast_div(double n, double d) {
  const double y = fast_rcp(d);
  const double q = n * y;
  return fma(fma(-d, q, n), y, q);      // one correction step: within 1 ulp
}
__device__ __forceinline__ double fast_sqrt(double a) {
  const double y = (double)rsqrtf((float)a);
  double g = a * y, h = 0.5 * y;
  double r = fma(-g, h, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  r = fma(-g, h, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  return fma(fma(-g, g, a), h, g);
}

// One level's contribution to the 32 degrees this thread owns:  acc[j] += u * r^j.
// Stride-8 scheme: 7 powers + per stride one add and seven FMAs
// (~1.25 fp64-pipe slots per degree instead of 2 for "acc += t; t *= r").
__device__ __forceinline__ void accumulate_level(double (&acc)[32], double u, double r, int half) {
  const double r2 = r * r, r3 = r2 * r, r4 = r2 * r2;
  const double r5 = r4 * r, r6 = r4 * r2, r7 = r4 * r3, r8 = r4 * r4;
  const double r16 = r8 * r8, r32 = r16 * r16;
  double us[4];
  us[0] = u * (half ? r32 : 1.0);
  us[1] = us[0] * r8;
  us[2] = us[0] * r16;
  us[3] = us[1] * r16;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const double v = us[s];
    acc[8 * s + 0] += v;
    acc[8 * s + 1] = fma(v, r, acc[8 * s + 1]);
    acc[8 * s + 2] = fma(v, r2, acc[8 * s + 2]);
    acc[8 * s + 3] = fma(v, r3, acc[8 * s + 3]);
    acc[8 * s + 4] = fma(v, r4, acc[8 * s + 4]);
    acc[8 * s + 5] = fma(v, r5, acc[8 * s + 5]);
    acc[8 * s + 6] = fma(v, r6, acc[8 * s + 6]);
    acc[8 * s + 7] = fma(v, r7, acc[8 * s + 7]);
  }
}

// per-column constants of the atmosphere geometry
struct ColumnGeom {
  double NG, NG1, cphi, sphi, geo_over;
};

// (u0, r) of one (level, column) element: u0 = -(dp/gamma) r^2 (BC05) or -(dp/g0) r^4 (SW02)
template <int MODE>
__device__ __forceinline__ double2 level_element(double z, double dpv, const ColumnGeom& c, double a1, double a2,
                                                 double st, double ct, double gs, double c1, double rad_e,
                                                 double inv_rad) {
  double u0, r;
  if (MODE == MODE_ATM_BC05) {
    const double H = a1 * z + a2 * (z * z);                 // orthometric height (List 1958), :222-224
    const double A = c.NG + H, B = c.NG1 + H;
    const double As = A * st;
    const double X = As * c.cphi, Y = As * c.sphi, Z = B * ct;
    const double R = fast_sqrt(fma(Z, Z, fma(Y, Y, X * X))); // :226-230
    r = R * inv_rad;
    const double hr = H * inv_rad;
    const double gam = gs * fma(3.0 * hr, hr, fma(-c1, hr, 1.0));   // :232-238
    const double q = fast_div(dpv, gam);
    u0 = -(q * r) * r;
  } else {
    r = fast_div(rad_e, rad_e - z) + c.geo_over;            // :258-260
    const double q = fast_div(dpv, 9.80665);
    const double r2 = r * r;
    u0 = -(q * r2) * r2;
  }
  return make_double2(u0, r);
}

// ---------------------------------------------------------------------------------------
// Produce phase, atmosphere (gen_atmosphere_stokes.py:217-265 fused):
//   B[p,l] = - sum_k (dp_k/gamma_k) (R_k/rad_e)^(l+2)                      BC05
//   B[p,l] = - sum_k (dp_k/g0) (rad_e/(rad_e-GPH_k) + GEOID/rad_e)^(l+4)   SW02
// (the minus sign is the reference's einsum(m_phi, -pfactor), :270).
// Two passes per group of LEV_GROUP levels:
//   geometry pass   : elementwise over (level, column), coalesced along lon, 4 independent
//                     elements in flight per thread; (u0, r) staged in shared memory
//   accumulate pass : lane pair (lane, lane^16) shares a column and splits the 64 degrees;
//                     32 running sums per thread in registers
// ---------------------------------------------------------------------------------------
template <int MODE, typename Tin>
__device__ __forceinline__ void produce_atmosphere(const GridParams& p, int t, int h, int lb, int chunk,
                                                   double* Bs) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double2* stg = reinterpret_cast<double2*>(Bs);            // [level in group][128 columns]
  const size_t lev_stride = (size_t)p.nlat * p.nlon;
  const int nlev = p.nlev;
  const double inv_rad = 1.0 / p.rad_e;
  double sg = 1.0, sp = 1.0;
  if (p.field_scale != nullptr) {
    sg = p.field_scale[2 * t];
    sp = p.field_scale[2 * t + 1];
  }
  // ring (colatitude-only) factors, built on the host
  const double a1 = p.ring_tab[0 * p.nlat + h], a2 = p.ring_tab[1 * p.nlat + h];
  const double st = p.ring_tab[4 * p.nlat + h], ct = p.ring_tab[5 * p.nlat + h];
  const double gs = p.ring_tab[6 * p.nlat + h], c1 = p.ring_tab[7 * p.nlat + h];

  // geometry-pass ownership: a fixed column, levels of parity tid>>7
  const int gcol = tid & (KC_MAX - 1), gpar = tid >> 7;
  const int gpcol = chunk * p.kc + gcol;
  const bool gactive = (gcol < p.kc) && (gpcol < p.nlon);
  const int gpc = gactive ? gpcol : 0;
  ColumnGeom cg;
  {
    const double geo = (p.geoid != nullptr) ? __ldg(p.geoid + h * p.geo_s0 + gpc * p.geo_s1) : 0.0;
    cg.NG = p.ring_tab[2 * p.nlat + h] + geo;     // (N + GEOID)
    cg.NG1 = p.ring_tab[3 * p.nlat + h] + geo;    // (N(1-e^2) + GEOID)
    cg.cphi = p.lon_tab[gpc];
    cg.sphi = p.lon_tab[p.nlon + gpc];
    cg.geo_over = geo / p.rad_e;
  }
  const Tin* zp = (const Tin*)p.GPH + (size_t)t * p.batch_stride + (size_t)h * p.nlon + gpc;
  const Tin* dp = (const Tin*)p.dP + (size_t)t * p.batch_stride + (size_t)h * p.nlon + gpc;

  // accumulate-pass ownership
  const int acol = warp * 16 + (lane & 15), half = lane >> 4;
  double acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.0;

  for (int kg = 0; kg < nlev; kg += LEV_GROUP) {
    const int ng = min(LEV_GROUP, nlev - kg);
    // ---- geometry pass (loads of the next 4 elements are in flight while 4 are computed)
    double zc[4], dc[4];
    auto load4 = [&](double(&z)[4], double(&d)[4], int kk0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kk = kk0 + 2 * i;
        z[i] = 0.0;
        d[i] = 0.0;
        if (gactive && kk < ng) {
          z[i] = ld_as_double(zp + (size_t)(kg + kk) * lev_stride);
          d[i] = ld_as_double(dp + (size_t)(kg + kk) * lev_stride);
        }
      }
    };
    load4(zc, dc, gpar);
    for (int kk0 = gpar; kk0 < ng; kk0 += 8) {
      double zn[4], dn[4];
      load4(zn, dn, kk0 + 8);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kk = kk0 + 2 * i;
        double2 e = level_element<MODE>(sg * zc[i], sp * dc[i], cg, a1, a2, st, ct, gs, c1, p.rad_e, inv_rad);
        if (lb > 0) e.x *= ipow(e.y, lb * LBLK);
        if (!gactive) e = make_double2(0.0, 1.0);
        if (kk < ng) stg[kk * KC_MAX + gcol] = e;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        zc[i] = zn[i];
        dc[i] = dn[i];
      }
    }
    __syncthreads();
    // ---- accumulate pass
#pragma unroll 2
    for (int kk = 0; kk < ng; ++kk) {
      const double2 e = stg[kk * KC_MAX + acol];
      accumulate_level(acc, e.x, e.y, half);
    }
    __syncthreads();      // staging is overwritten by the next group / by the B chunk
  }
  if (acol < p.kc) {
    const int l0 = 32 * half;
#pragma unroll
    for (int j = 0; j < 32; ++j) Bs[bs_index(l0 + j, acol)] = acc[j];
  }
}

// L2 prefetch of the cube data the next chunk's geometry pass will read (issued before the
// contraction phase so that DRAM latency is off the critical path)
template <typename Tin>
__device__ __forceinline__ void prefetch_chunk(const GridParams& p, int t, int h, int chunk) {
  if (h >= p.nlat) return;
  const size_t lev_stride = (size_t)p.nlat * p.nlon;
  const int per_line = 128 / (int)sizeof(Tin);
  const int lines = (p.kc + per_line - 1) / per_line + 1;
  const size_t base = (size_t)t * p.batch_stride + (size_t)h * p.nlon + (size_t)chunk * p.kc;
  for (int i = threadIdx.x; i < p.nlev * lines * 2; i += NT) {
    const int arr = i & 1, j = i >> 1;
    const int k = j / lines, ln = j % lines;
    const int col = min(ln * per_line, p.kc - 1);
    if ((size_t)chunk * p.kc + col >= (size_t)p.nlon) continue;
    const Tin* q = (const Tin*)(arr ? p.dP : p.GPH) + base + (size_t)k * lev_stride + col;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
  }
}

// ---------------------------------------------------------------------------------------
// One segment: work units [u0, u1) (unit = ring * nchunks + chunk) of tile (lb, mb) of field t.
//
// DIAG tile (lb == mb): only the 8x8 sub-tiles with degree-tile >= order-tile are needed (36 of
// 64).  The four warps {q, q+4} ... share the order groups gA = q and gB = 7-q, i.e. 9
// (order-group, degree-tile) pairs j = 0..8.  Warp q owns pairs 0..3, warp q+4 owns pairs 5..8,
// and the "swing" pair 4 is done by warp q on even k-steps and by warp q+4 on odd k-steps:
// 9 DMMA pairs per two k-steps for every warp (perfectly balanced), 5 accumulator pairs per
// thread.  FULL tile: warp w owns order group w and all 8 degree tiles.
// ---------------------------------------------------------------------------------------
#ifndef MHB_SWING
#define MHB_SWING 0
#endif
template <int MODE, typename Tin, bool DIAG>
__device__ __forceinline__ void run_segment(const GridParams& p, const Segment seg, int slot_index,
                                            double* Bs, double* Cs, double* plm_s) {
  // SW = swing mapping (above); !SW = each warp of a quad does all 9 pairs on k-steps of one parity
  constexpr bool SW = (MHB_SWING != 0);
  constexpr int NP = DIAG ? (SW ? 5 : 9) : 8;          // accumulator pairs per thread
  constexpr int KSTRIDE = (DIAG && !SW) ? 2 : 1;
  constexpr int NA = DIAG ? 4 : 2;          // trig fragments per k-step
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = DIAG ? (warp >> 2) : 0;
  const int gA = DIAG ? (warp & 3) : warp;
  const int gB = DIAG ? (7 - gA) : warp;
  const int nA = DIAG ? (8 - gA) : 8;
  const int nksc = p.kc >> 2;

  int lt[NP];          // degree tile of local pair j
  bool inA[NP];        // pair belongs to order group gA (else gB)
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    if (DIAG) {
      const int jg = SW ? ((j < 4) ? (sub ? 5 + j : j) : 4) : j;     // index in the quad's list of 9 pairs
      inA[j] = (jg < nA);
      lt[j] = inA[j] ? (gA + jg) : (jg - 1);
    } else {
      inA[j] = true;
      lt[j] = j;
    }
  }

#pragma unroll
  for (int i = 0; i < NP * 4; ++i) Cs[i * NT + tid] = 0.0;

  const int hfirst = seg.u0 / p.nchunks, hlast = (seg.u1 - 1) / p.nchunks;
  for (int h = hfirst; h <= hlast; ++h) {
    // chunks of this ring that belong to the segment (partial rings are fine: the Legendre
    // weighting below is linear in the longitude partial sums)
    const int cbeg = (h == hfirst) ? (seg.u0 - h * p.nchunks) : 0;
    const int cend = (h == hlast) ? (seg.u1 - h * p.nchunks) : p.nchunks;
    if (tid < MBLK) legendre_ring(p, h, seg.lb, seg.mb, tid, plm_s);

    double acc[NP][2][2];
#pragma unroll
    for (int j = 0; j < NP; ++j) acc[j][0][0] = acc[j][0][1] = acc[j][1][0] = acc[j][1][1] = 0.0;

    for (int chunk = cbeg; chunk < cend; ++chunk) {
      if (MODE == MODE_PRESSURE)
        produce_pressure(p, seg.t, h, seg.lb, chunk, Bs);
      else
        produce_atmosphere<MODE, Tin>(p, seg.t, h, seg.lb, chunk, Bs);
      __syncthreads();

      if (MODE != MODE_PRESSURE) {
        // start pulling the next chunk's (or next ring's first chunk's) cube data into L2
        const bool last = (chunk + 1 == p.nchunks);
        prefetch_chunk<Tin>(p, seg.t, last ? h + 1 : h, last ? 0 : chunk + 1);
      }
      // contraction over the chunk's longitudes; trig fragments straight from L2 (fragment-major
      // table: one coalesced 256-byte load per fragment, prefetched three k-steps ahead in a
      // register ring), B fragments from shared memory
      const double* Ablk = p.trig + ((size_t)(seg.mb * p.nks_total + chunk * nksc) * 16) * 32 + lane;
      auto load_a = [&](double(&a)[NA], int k) {
        if (k < nksc) {
          const double* ap = Ablk + (size_t)k * 512;
          a[0] = __ldg(ap + (gA * 2) * 32);
          a[1] = __ldg(ap + (gA * 2 + 1) * 32);
          if (DIAG && (sub || !SW)) {        // swing: only the second warp of a quad touches order group gB
            a[2] = __ldg(ap + (gB * 2) * 32);
            a[3] = __ldg(ap + (gB * 2 + 1) * 32);
          }
        }
      };
      double a[4][NA];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NA; ++j) a[i][j] = 0.0;
      const int k0 = (DIAG && !SW) ? sub : 0;
#pragma unroll
      for (int i = 0; i < 3; ++i) load_a(a[i], k0 + i * KSTRIDE);
      // 4-deep ring, fully unrolled so no rotation copies shorten the prefetch distance;
      // with stride 1 the k-step parity of ring slot u is u & 1 (compile time)
      for (int ks = k0; ks < nksc; ks += 4 * KSTRIDE) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int k = ks + u * KSTRIDE;
          if (k < nksc) {
            load_a(a[(u + 3) & 3], k + 3 * KSTRIDE);
            const double(&c)[NA] = a[u];
            const double* Bk = Bs + k * 256 + ((((lane >> 2) + k) & 7) << 2) + (lane & 3);
            double b[NP];
#pragma unroll
            for (int j = 0; j < NP; ++j) b[j] = Bk[lt[j] * 32];
#pragma unroll
            for (int j = 0; j < NP; ++j) {
              if (DIAG && SW && j == 4 && ((u & 1) != sub)) continue;     // swing pair: other warp's turn
              dmma884(acc[j][0][0], acc[j][0][1], inA[j] ? c[0] : c[DIAG ? 2 : 0], b[j]);
              dmma884(acc[j][1][0], acc[j][1][1], inA[j] ? c[1] : c[DIAG ? 3 : 1], b[j]);
            }
          }
        }
      }
      __syncthreads();
    }

    // ring epilogue: C += (w_h Plm) o Dring
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const int g = inA[j] ? gA : gB;
      const int m_local = g * 8 + (lane >> 2);
      const int l_local = lt[j] * 8 + 2 * (lane & 3);
      const double w0 = plm_s[m_local * PLM_PITCH + l_local];
      const double w1 = plm_s[m_local * PLM_PITCH + l_local + 1];
#pragma unroll
      for (int cs = 0; cs < 2; ++cs) {
        double* c = Cs + ((j * 2 + cs) * 2) * NT + tid;
        c[0] = fma(w0, acc[j][cs][0], c[0]);
        c[NT] = fma(w1, acc[j][cs][1], c[NT]);
      }
    }
    __syncthreads();
  }

  // write the segment's partial sums: slot[cs][l_local][m_local]
  double* slot = p.slots + (size_t)slot_index * SLOT_DOUBLES;
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    if (DIAG && SW && j == 4 && sub) continue;        // the swing pair is written by the first warp
    if (DIAG && !SW && sub) continue;             // parity mapping: first warp writes the sum of both
    const int g = inA[j] ? gA : gB;
    const int m_local = g * 8 + (lane >> 2);
    const int l_local = lt[j] * 8 + 2 * (lane & 3);
#pragma unroll
    for (int cs = 0; cs < 2; ++cs) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int row = ((j * 2 + cs) * 2 + e) * NT;
        double v = Cs[row + tid];
        if (DIAG && (j == 4 || !SW)) v += Cs[row + tid + 128];   // even + odd k-steps
        slot[(cs * LBLK + l_local + e) * MBLK + m_local] = v;
      }
    }
  }
  __syncthreads();
}

template <int MODE, typename Tin>
__global__ void __launch_bounds__(NT, 1) stokes_ring_kernel(const GridParams p) {
  extern __shared__ double smem[];
  double* Bs = smem;
  double* Cs = Bs + REGION0_DOUBLES;
  double* plm_s = Cs + CS_ROWS * NT;
  const int sb = p.cta_seg_begin[blockIdx.x], se = p.cta_seg_begin[blockIdx.x + 1];
  for (int s = sb; s < se; ++s) {
    const Segment seg = p.segs[s];
    if (seg.lb == seg.mb)
      run_segment<MODE, Tin, true>(p, seg, s, Bs, Cs, plm_s);
    else
      run_segment<MODE, Tin, false>(p, seg, s, Bs, Cs, plm_s);
  }
}

// Fixed-order reduction of the partial slots + degree factors (gen_pressure_stokes.py:205-206).
// grid (lmax+1, nbatch), 256 threads: 4 interleaved partial sums per (l, m) -- independent
// loads in flight -- combined in a fixed order, so results are bit-reproducible.
__global__ void __launch_bounds__(256) stokes_reduce_kernel(const __grid_constant__ ReduceParams q) {
  __shared__ double part[2][4][MBLK];
  const int l = blockIdx.x, t = blockIdx.y;
  const int ncol = q.mmax + 1;
  const double df = (q.dfactor_dev != nullptr) ? q.dfactor_dev[l] : q.dfactor[l];
  const int lb = l / LBLK, l_local = l % LBLK;
  const int ml = threadIdx.x & (MBLK - 1), grp = threadIdx.x >> 6;
  for (int mb = 0; mb * MBLK < ncol; ++mb) {
    const int m = mb * MBLK + ml;
    double c = 0.0, s = 0.0;
    if (m <= l && m < ncol) {
      const int tile = t * q.ntiles + q.lb_first_tile[lb] + mb;
      const int b = q.tile_slot_begin[tile], e = q.tile_slot_begin[tile + 1];
#pragma unroll 4
      for (int k = b + grp; k < e; k += 4) {
        const double* slot = q.slots + (size_t)k * SLOT_DOUBLES;
        c += slot[(l_local)*MBLK + ml];
        s += slot[(LBLK + l_local) * MBLK + ml];
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

// =========================================================================================
// host side: plan, schedule, entry points
// =========================================================================================
using namespace mhb;

struct Schedule {
  int nbatch = -1, nlev = -1, mode = -1;
  int ncta = 0, nseg = 0;
  Segment* segs_dev = nullptr;
  int* cta_seg_begin_dev = nullptr;
  int* tile_slot_begin_dev = nullptr;
};

struct mhb200_plan {
  int device, nlat, nlon, lmax, mmax;
  int nlb, nmb, Mpad, ntiles;
  int kc, nchunks, nks_total, num_sms;
  double *trig, *f1, *f2, *pmm, *sq, *rescale, *x, *w;
  int* lb_first_tile_dev;
  std::vector<int> lb_first_tile;
  // per-call staging (dfactor, field scales), grown on demand
  double* stage_dev;
  size_t stage_doubles;
  size_t stage_reserved;   // doubles at the front of the staging buffer used by field scales
  std::mutex mu;
  Schedule sched;
};

namespace {

template <typename T>
int upload(T** dev, const std::vector<T>& host) {
  MHB_CUDA(cudaMalloc((void**)dev, std::max<size_t>(host.size(), 1) * sizeof(T)));
  if (!host.empty()) MHB_CUDA(cudaMemcpy(*dev, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
  return MHB200_OK;
}

// fp64-pipe cost (arbitrary units per longitude) of one ring of one tile
double ring_cost(bool diag, int mode, int nlev) {
  const double mma = diag ? 4608.0 : 8192.0;
  const double prod = (mode == MODE_PRESSURE) ? 90.0 : (double)nlev * 136.0;
  return mma + prod;
}

int build_schedule(mhb200_plan* pl, int nbatch, int nlev, int mode) {
  Schedule& sc = pl->sched;
  if (sc.nbatch == nbatch && sc.nlev == nlev && sc.mode == mode) return MHB200_OK;
  if (sc.segs_dev) cudaFree(sc.segs_dev);
  if (sc.cta_seg_begin_dev) cudaFree(sc.cta_seg_begin_dev);
  if (sc.tile_slot_begin_dev) cudaFree(sc.tile_slot_begin_dev);
  sc = Schedule();
  // linearised work: (field, tile, ring) with per-ring cost
  struct Tile { int lb, mb; double c; };
  std::vector<Tile> tiles;
  for (int lb = 0; lb < pl->nlb; ++lb)
    for (int mb = 0; mb <= std::min(lb, pl->nmb - 1); ++mb)
      tiles.push_back({lb, mb, ring_cost(lb == mb, mode, nlev)});
  // per-unit cost = per-ring cost / nchunks (only ratios matter)
  const int upr = pl->nchunks;                       // units per ring
  const long long units_per_tile = (long long)pl->nlat * upr;
  double per_field = 0.0;
  for (auto& tl : tiles) per_field += tl.c * units_per_tile;
  const double total = per_field * nbatch;
  const long long units = (long long)nbatch * (long long)tiles.size() * units_per_tile;
  const int ncta = (int)std::min<long long>(pl->num_sms, units);
  std::vector<Segment> segs;
  std::vector<int> cta_begin(ncta + 1, 0);
  std::vector<int> tile_slot_begin((size_t)nbatch * tiles.size() + 1, 0);
  double done = 0.0;
  int cta = 0;
  for (int t = 0; t < nbatch; ++t) {
    for (size_t ti = 0; ti < tiles.size(); ++ti) {
      tile_slot_begin[(size_t)t * tiles.size() + ti] = (int)segs.size();
      const double c = tiles[ti].c;
      long long u = 0;
      while (u < units_per_tile) {
        // units the current CTA can still take before its cost boundary
        const double boundary = total * (double)(cta + 1) / (double)ncta;
        long long take = (long long)std::floor((boundary - done) / c + 0.5);
        if (cta == ncta - 1) take = units_per_tile - u;
        take = std::max<long long>(0, std::min<long long>(take, units_per_tile - u));
        if (take > 0) {
          segs.push_back({t, tiles[ti].lb, tiles[ti].mb, (int)u, (int)(u + take), 0});
          u += take;
          done += take * c;
        }
        if (u < units_per_tile && cta < ncta - 1) {   // this CTA is full: move to the next one
          ++cta;
          cta_begin[cta] = (int)segs.size();
        }
      }
      // tile finished exactly at (or past) the CTA's boundary: next tile starts on a new CTA
      if (cta < ncta - 1 && done >= total * (double)(cta + 1) / (double)ncta - 0.5 * c) {
        ++cta;
        cta_begin[cta] = (int)segs.size();
      }
    }
  }
  for (int c = cta + 1; c <= ncta; ++c) cta_begin[c] = (int)segs.size();
  tile_slot_begin.back() = (int)segs.size();
  sc.ncta = ncta;
  sc.nseg = (int)segs.size();
  int rc;
  if ((rc = upload(&sc.segs_dev, segs)) != MHB200_OK) return rc;
  if ((rc = upload(&sc.cta_seg_begin_dev, cta_begin)) != MHB200_OK) return rc;
  if ((rc = upload(&sc.tile_slot_begin_dev, tile_slot_begin)) != MHB200_OK) return rc;
  sc.nbatch = nbatch;
  sc.nlev = nlev;
  sc.mode = mode;
  return MHB200_OK;
}

size_t max_segments(const mhb200_plan* pl, int nbatch) {
  // every (field, tile) contributes at least one segment and every CTA boundary at most one more
  return (size_t)nbatch * pl->ntiles + pl->num_sms + 1;
}

int stage(mhb200_plan* pl, const double* host, size_t n, size_t offset, cudaStream_t st) {
  if (offset + n > pl->stage_doubles) {
    set_err("internal: staging buffer too small");
    return MHB200_EINVAL;
  }
  MHB_CUDA(cudaMemcpyAsync(pl->stage_dev + offset, host, n * sizeof(double), cudaMemcpyHostToDevice, st));
  return MHB200_OK;
}

int ensure_stage(mhb200_plan* pl, size_t n) {
  if (n <= pl->stage_doubles) return MHB200_OK;
  if (pl->stage_dev) {
    MHB_CUDA(cudaDeviceSynchronize());
    cudaFree(pl->stage_dev);
    pl->stage_dev = nullptr;
    pl->stage_doubles = 0;
  }
  MHB_CUDA(cudaMalloc((void**)&pl->stage_dev, n * sizeof(double)));
  pl->stage_doubles = n;
  return MHB200_OK;
}

template <int MODE, typename Tin>
int launch_ring(const GridParams& gp, int ncta, cudaStream_t st) {
  auto kern = stokes_ring_kernel<MODE, Tin>;
  MHB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  kern<<<ncta, NT, SMEM_BYTES, st>>>(gp);
  g_launches.fetch_add(1);
  MHB_CUDA(cudaGetLastError());
  return MHB200_OK;
}

int finish_reduce(mhb200_plan* pl, const double* dfactor, int nbatch, double* slots, double* clm, double* slm,
                  cudaStream_t st) {
  ReduceParams q;
  q.dfactor_dev = nullptr;
  if (pl->lmax + 1 <= DF_MAX) {
    for (int l = 0; l <= pl->lmax; ++l) q.dfactor[l] = dfactor[l];
  } else {
    int rc;
    if ((rc = ensure_stage(pl, (size_t)pl->lmax + 1 + pl->stage_reserved)) != MHB200_OK) return rc;
    if ((rc = stage(pl, dfactor, (size_t)pl->lmax + 1, pl->stage_reserved, st)) != MHB200_OK) return rc;
    q.dfactor_dev = pl->stage_dev + pl->stage_reserved;
  }
  q.lmax = pl->lmax;
  q.mmax = pl->mmax;
  q.nmb = pl->nmb;
  q.ntiles = pl->ntiles;
  q.tile_slot_begin = pl->sched.tile_slot_begin_dev;
  q.lb_first_tile = pl->lb_first_tile_dev;
  q.slots = slots;
  q.clm = clm;
  q.slm = slm;
  dim3 grid(pl->lmax + 1, nbatch);
  stokes_reduce_kernel<<<grid, 256, 0, st>>>(q);
  g_launches.fetch_add(1);
  MHB_CUDA(cudaGetLastError());
  return MHB200_OK;
}

void fill_plan_params(const mhb200_plan* pl, GridParams& gp) {
  memset(&gp, 0, sizeof(gp));
  gp.nlat = pl->nlat;
  gp.nlon = pl->nlon;
  gp.lmax = pl->lmax;
  gp.mmax = pl->mmax;
  gp.Mpad = pl->Mpad;
  gp.kc = pl->kc;
  gp.nchunks = pl->nchunks;
  gp.nks_total = pl->nks_total;
  gp.trig = pl->trig;
  gp.f1 = pl->f1;
  gp.f2 = pl->f2;
  gp.pmm = pl->pmm;
  gp.sq = pl->sq;
  gp.rescale = pl->rescale;
  gp.x = pl->x;
  gp.w = pl->w;
  gp.segs = pl->sched.segs_dev;
  gp.cta_seg_begin = pl->sched.cta_seg_begin_dev;
}

}  // namespace

extern "C" {

int mhb200_plan_create(mhb200_plan** out, int device, int nlat, int nlon, int lmax, int mmax,
                       const double* phi, const double* costh, const double* int_fact) {
  if (!out || nlat < 1 || nlon < 1 || lmax < 0 || mmax < 0 || mmax > lmax || !phi || !costh || !int_fact) {
    set_err("mhb200_plan_create: bad argument");
    return MHB200_EINVAL;
  }
  int rc = mhb200_device_check(device);
  if (rc != MHB200_OK) return rc;
  MHB_CUDA(cudaSetDevice(device));
  mhb200_plan* pl = new mhb200_plan();
  pl->device = device;
  pl->nlat = nlat;
  pl->nlon = nlon;
  pl->lmax = lmax;
  pl->mmax = mmax;
  pl->nlb = (lmax + LBLK) / LBLK;
  pl->nmb = (mmax + MBLK) / MBLK;
  pl->Mpad = pl->nmb * MBLK;
  pl->nchunks = (nlon + KC_MAX - 1) / KC_MAX;
  pl->kc = (((nlon + pl->nchunks - 1) / pl->nchunks) + 7) / 8 * 8;
  pl->nks_total = pl->nchunks * (pl->kc / 4);
  pl->stage_dev = nullptr;
  pl->stage_doubles = 0;
  pl->stage_reserved = 0;
  cudaDeviceProp prop;
  MHB_CUDA(cudaGetDeviceProperties(&prop, device));
  pl->num_sms = prop.multiProcessorCount;
  pl->lb_first_tile.resize(pl->nlb);
  int nt = 0;
  for (int lb = 0; lb < pl->nlb; ++lb) {
    pl->lb_first_tile[lb] = nt;
    nt += std::min(lb + 1, pl->nmb);
  }
  pl->ntiles = nt;

  // trig table, fragment-major [mb][kstep][16 row tiles][32 lanes]  (m_phi, gen_pressure_stokes.py:176-177)
  const int npad = pl->nks_total * 4;
  std::vector<double> trig((size_t)pl->nmb * pl->nks_total * 512, 0.0);
  for (int mb = 0; mb < pl->nmb; ++mb)
    for (int ks = 0; ks < pl->nks_total; ++ks)
      for (int g = 0; g < 8; ++g)
        for (int lane = 0; lane < 32; ++lane) {
          const int m = mb * MBLK + g * 8 + (lane >> 2);
          const int p = ks * 4 + (lane & 3);
          if (m > mmax || p >= nlon || p >= npad) continue;
          const double arg = (double)m * phi[p];
          const size_t base = (((size_t)mb * pl->nks_total + ks) * 16 + g * 2) * 32 + lane;
          trig[base] = cos(arg);
          trig[base + 32] = sin(arg);
        }
  // Holmes-Featherstone multipliers [(lmax+1)][Mpad], sectoral seeds and sqrt(2m+3)
  const int Mpad = pl->Mpad;
  std::vector<double> f1((size_t)(lmax + 1) * Mpad, 0.0), f2((size_t)(lmax + 1) * Mpad, 0.0);
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
  const double scalef = 1.0e-280;
  std::vector<double> pmm(Mpad, 0.0), sq(Mpad, 0.0);
  pmm[0] = 1.0;   // zonal column is unscaled
  {
    double v = sqrt(2.0) * scalef;
    for (int m = 1; m <= mmax; ++m) {
      v = v * sqrt(2.0 * m + 1.0) / sqrt(2.0 * m);
      pmm[m] = v;
    }
    for (int m = 0; m <= mmax; ++m) sq[m] = sqrt(2.0 * m + 3.0);
  }
  // rescale[h][m] = u^m / scalef by sequential products, u = sqrt(1-x^2) with 0 -> eps
  std::vector<double> rescale((size_t)nlat * Mpad, 0.0), xs(costh, costh + nlat), ws(int_fact, int_fact + nlat);
  for (int h = 0; h < nlat; ++h) {
    double u = sqrt(1.0 - costh[h] * costh[h]);
    if (u == 0.0) u = 2.220446049250313e-16;
    double r = 1.0 / scalef;
    rescale[(size_t)h * Mpad] = 1.0;
    for (int m = 1; m <= mmax; ++m) {
      r = r * u;
      rescale[(size_t)h * Mpad + m] = r;
    }
  }
  if ((rc = upload(&pl->trig, trig)) || (rc = upload(&pl->f1, f1)) || (rc = upload(&pl->f2, f2)) ||
      (rc = upload(&pl->pmm, pmm)) || (rc = upload(&pl->sq, sq)) || (rc = upload(&pl->rescale, rescale)) ||
      (rc = upload(&pl->x, xs)) || (rc = upload(&pl->w, ws)) ||
      (rc = upload(&pl->lb_first_tile_dev, pl->lb_first_tile))) {
    delete pl;
    return rc;
  }
  *out = pl;
  return MHB200_OK;
}

int mhb200_plan_destroy(mhb200_plan* pl) {
  if (!pl) return MHB200_OK;
  cudaSetDevice(pl->device);
  cudaFree(pl->trig);
  cudaFree(pl->f1);
  cudaFree(pl->f2);
  cudaFree(pl->pmm);
  cudaFree(pl->sq);
  cudaFree(pl->rescale);
  cudaFree(pl->x);
  cudaFree(pl->w);
  cudaFree(pl->lb_first_tile_dev);
  if (pl->stage_dev) cudaFree(pl->stage_dev);
  if (pl->sched.segs_dev) cudaFree(pl->sched.segs_dev);
  if (pl->sched.cta_seg_begin_dev) cudaFree(pl->sched.cta_seg_begin_dev);
  if (pl->sched.tile_slot_begin_dev) cudaFree(pl->sched.tile_slot_begin_dev);
  delete pl;
  return MHB200_OK;
}

size_t mhb200_grid_workspace_bytes(const mhb200_plan* pl, int nbatch, int nlev) {
  (void)nlev;
  if (!pl || nbatch < 1) return 0;
  return max_segments(pl, nbatch) * SLOT_DOUBLES * sizeof(double);
}

int mhb200_pressure_stokes_f64(const mhb200_plan* cpl, const double* P, const int64_t Ps[3], const double* G,
                               const int64_t Gs[3], const double* R, const int64_t Rs[3], int nbatch,
                               double rad_e, const double* dfactor, const double* plm_table, double* clm_out,
                               double* slm_out, void* workspace, size_t workspace_bytes, void* stream) {
  mhb200_plan* pl = const_cast<mhb200_plan*>(cpl);
  if (!pl || !P || !G || !R || !Ps || !Gs || !Rs || nbatch < 1 || !dfactor || !clm_out || !slm_out || !workspace) {
    set_err("mhb200_pressure_stokes_f64: bad argument");
    return MHB200_EINVAL;
  }
  if (workspace_bytes < mhb200_grid_workspace_bytes(pl, nbatch, 0)) {
    set_err("mhb200_pressure_stokes_f64: workspace too small");
    return MHB200_EWORKSPACE;
  }
  std::lock_guard<std::mutex> lock(pl->mu);
  cudaStream_t st = (cudaStream_t)stream;
  MHB_CUDA(cudaSetDevice(pl->device));
  int rc;
  if ((rc = build_schedule(pl, nbatch, 0, MODE_PRESSURE)) != MHB200_OK) return rc;
  pl->stage_reserved = 0;
  GridParams gp;
  fill_plan_params(pl, gp);
  gp.plm_table = plm_table;
  gp.slots = (double*)workspace;
  gp.P = P;
  gp.G = G;
  gp.R = R;
  for (int i = 0; i < 3; ++i) {
    gp.Ps[i] = Ps[i];
    gp.Gs[i] = Gs[i];
    gp.Rs[i] = Rs[i];
  }
  gp.rad_e = rad_e;
  if ((rc = launch_ring<MODE_PRESSURE, double>(gp, pl->sched.ncta, st)) != MHB200_OK) return rc;
  return finish_reduce(pl, dfactor, nbatch, gp.slots, clm_out, slm_out, st);
}

int mhb200_atmosphere_stokes(const mhb200_plan* cpl, int dtype, const void* GPH, const void* dP,
                             int64_t batch_stride, const double* field_scale, const double* geoid,
                             const int64_t geoid_stride[2], const double* ring_tab, const double* lon_tab,
                             int nlev, int nbatch, int method, double rad_e, const double* dfactor,
                             const double* plm_table, double* clm_out, double* slm_out, void* workspace,
                             size_t workspace_bytes, void* stream) {
  mhb200_plan* pl = const_cast<mhb200_plan*>(cpl);
  if (!pl || !GPH || !dP || !ring_tab || !lon_tab || nlev < 1 || nbatch < 1 || !dfactor || !clm_out ||
      !slm_out || !workspace || (dtype != MHB200_F64 && dtype != MHB200_F32) ||
      (method != MHB200_METHOD_BC05 && method != MHB200_METHOD_SW02) || (geoid && !geoid_stride)) {
    set_err("mhb200_atmosphere_stokes: bad argument");
    return MHB200_EINVAL;
  }
  if (workspace_bytes < mhb200_grid_workspace_bytes(pl, nbatch, nlev)) {
    set_err("mhb200_atmosphere_stokes: workspace too small");
    return MHB200_EWORKSPACE;
  }
  std::lock_guard<std::mutex> lock(pl->mu);
  cudaStream_t st = (cudaStream_t)stream;
  MHB_CUDA(cudaSetDevice(pl->device));
  const int mode = (method == MHB200_METHOD_BC05) ? MODE_ATM_BC05 : MODE_ATM_SW02;
  int rc;
  if ((rc = build_schedule(pl, nbatch, nlev, mode)) != MHB200_OK) return rc;
  const size_t nl = (size_t)pl->lmax + 1;
  pl->stage_reserved = field_scale ? 2 * (size_t)nbatch : 0;
  if ((rc = ensure_stage(pl, nl + 2 * (size_t)nbatch)) != MHB200_OK) return rc;
  if (field_scale && (rc = stage(pl, field_scale, 2 * (size_t)nbatch, 0, st)) != MHB200_OK) return rc;
  GridParams gp;
  fill_plan_params(pl, gp);
  gp.plm_table = plm_table;
  gp.slots = (double*)workspace;
  gp.GPH = GPH;
  gp.dP = dP;
  gp.batch_stride = batch_stride;
  gp.field_scale = field_scale ? pl->stage_dev : nullptr;
  gp.geoid = geoid;
  gp.geo_s0 = geoid ? geoid_stride[0] : 0;
  gp.geo_s1 = geoid ? geoid_stride[1] : 0;
  gp.ring_tab = ring_tab;
  gp.lon_tab = lon_tab;
  gp.nlev = nlev;
  gp.rad_e = rad_e;
  const int ncta = pl->sched.ncta;
  if (mode == MODE_ATM_BC05)
    rc = (dtype == MHB200_F64) ? launch_ring<MODE_ATM_BC05, double>(gp, ncta, st)
                               : launch_ring<MODE_ATM_BC05, float>(gp, ncta, st);
  else
    rc = (dtype == MHB200_F64) ? launch_ring<MODE_ATM_SW02, double>(gp, ncta, st)
                               : launch_ring<MODE_ATM_SW02, float>(gp, ncta, st);
  if (rc != MHB200_OK) return rc;
  return finish_reduce(pl, dfactor, nbatch, gp.slots, clm_out, slm_out, st);
}

}  // extern "C"
