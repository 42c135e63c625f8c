/*
 * mhb200.h -- C ABI of the B200-native fields -> Stokes hot path.
 *
 * Drop-in boundary for the three operators of tsutterley/model-harmonics
 * (paths relative to the reference tree):
 *   model_harmonics/gen_pressure_stokes.py:81-209    -> mhb200_pressure_stokes_f64
 *   model_harmonics/gen_atmosphere_stokes.py:86-278  -> mhb200_atmosphere_stokes
 *   model_harmonics/gen_point_pressure.py:64-194     -> mhb200_point_pressure_f64
 * The reference has no FFI: its boundary is those three Python functions
 * (model_harmonics/__init__.py:19-21).  The Python package model_harmonics_b200 keeps
 * their signatures and binds this library with ctypes.
 *
 * Conventions
 *   - every data pointer is a DEVICE pointer owned by the caller unless the
 *     parameter comment says "host";
 *   - all arithmetic is IEEE binary64; float32 is accepted only as an input
 *     storage type for the 3-D cubes (converted exactly on load);
 *   - outputs clm/slm are (nbatch, lmax+1, mmax+1) row-major, fully written
 *     (zeros where m > l);
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*); the
 *     calls do not synchronise;
 *   - re-entrancy: a plan is immutable after creation apart from internally locked caches, and every
 *     per-call mutable buffer (partial sums, staged field scales / degree factors, the moment path's
 *     Fourier sums and guard flag) lives in the CALLER's workspace.  Calls may therefore run
 *     concurrently from several threads and streams on one plan as long as each stream in flight
 *     has its own workspace (same-stream calls may share one).  The host-buffer entry
 *     mhb200_atmosphere_stokes_host owns device staging cubes per plan and is serialised per plan;
 *   - return value: MHB200_OK or a negative MHB200_E* code;
 *     mhb200_last_error() gives a thread-local message.
 *   - there is no CPU fallback: without a CUDA device every call fails.
 *
 * Kernel paths of the grid operators (same results to rounding)
 *   - contracted-moment path (default; csrc/stokes_moments.cu): regular longitude grids with both
 *     mirror symmetries, no caller-supplied PLM table; the atmosphere form also needs LMAX <= 63 and
 *     16-byte aligned cubes / rows (TMA).  It is guarded: every call also enqueues the direct kernels,
 *     which exit at once unless a radius left the validated range of the binomial expansion
 *     (|r / r0 - 1| <= 0.8 / (LMAX + 2 or 4)), in which case they recompute the call exactly.
 *   - direct kernels (csrc/stokes_grid.cu): any grid, any LMAX, caller PLM tables.
 *   MHB200_MOMENTS=0 forces the direct kernels, =1 the round-1 expanded-moment variant (unguarded).
 */
#ifndef MHB200_H_
#define MHB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MHB200_OK            0
#define MHB200_EINVAL       -1   /* bad argument */
#define MHB200_ECUDA        -2   /* CUDA runtime error (see mhb200_last_error) */
#define MHB200_EWORKSPACE   -3   /* workspace too small */
#define MHB200_ENODEVICE    -4   /* no usable CUDA device */

/* integration over pressure levels, gen_atmosphere_stokes.py:254-265 */
#define MHB200_METHOD_BC05   0
#define MHB200_METHOD_SW02   1

/* storage type of the (level, lat, lon) cubes */
#define MHB200_F64           0
#define MHB200_F32           1

typedef struct mhb200_plan mhb200_plan;

/* library / device probe; returns MHB200_OK when device `device` is an sm_100 part */
int mhb200_version(void);
int mhb200_device_check(int device);
const char* mhb200_last_error(void);

/*
 * Geometry plan: everything that depends only on (lon, lat, LMAX, MMAX, latitude
 * weights).  Replaces the tables the reference rebuilds per call:
 *   m_phi = exp(1j*m*phi)            gen_pressure_stokes.py:176-177
 *   PLM   = plm_holmes(LMAX,cos(th)) gen_pressure_stokes.py:180-182 (here: only the
 *           O(L^2) recursion multipliers and the O(nlat*M) u^m rescale factors are
 *           stored; P_lm itself is generated on the fly inside the kernel)
 *   plm*int_fact                     gen_pressure_stokes.py:186-188
 * phi[nlon]      host  longitudes in radians exactly as the operator forms them
 * costh[nlat]    host  cos(colatitude)
 * int_fact[nlat] host  latitude integration weights (WEIGHT or sin(th)*dphi*dth)
 */
int mhb200_plan_create(mhb200_plan** plan, int device, int nlat, int nlon, int lmax, int mmax,
                       const double* phi, const double* costh, const double* int_fact);
int mhb200_plan_destroy(mhb200_plan* plan);

/* bytes of device workspace a call with these sizes needs (0 on bad arguments) */
size_t mhb200_grid_workspace_bytes(const mhb200_plan* plan, int nbatch, int nlev);

/*
 * gen_pressure_stokes.py:194-206.  For each of nbatch fields:
 *   C_lm + i S_lm = dfactor[l] * sum_h w_h Plm(x_h) sum_p e^{i m phi_p} (P/G) (R/rad_e)^(l+2)
 * P, G, R are addressed as base[t*stride[0] + h*stride[1] + p*stride[2]] (element strides;
 * 0 broadcasts), which covers lat x lon, lon x lat, scalar and 1-element operands
 * (gen_pressure_stokes.py:146-151).
 * dfactor[lmax+1]  host
 * plm_table        NULL (on-the-fly Holmes-Featherstone recursion) or a device table
 *                  (nlat, lmax+1, mmax+1) row-major holding a caller-supplied PLM
 */
int mhb200_pressure_stokes_f64(const mhb200_plan* plan,
                               const double* P, const int64_t P_stride[3],
                               const double* G, const int64_t G_stride[3],
                               const double* R, const int64_t R_stride[3],
                               int nbatch, double rad_e, const double* dfactor,
                               const double* plm_table,
                               double* clm_out, double* slm_out,
                               void* workspace, size_t workspace_bytes, void* stream);

/*
 * gen_atmosphere_stokes.py:217-275, geometry pre-pass and level integration fused.
 * GPH, dP          (nbatch, nlev, nlat, nlon) cubes of type `dtype` (MHB200_F64/F32)
 * field_scale      host, NULL or 2*nbatch doubles (sg_t, sp_t): field t is
 *                  (sg_t*GPH, sp_t*dP) formed on load -- lets a batch share one base cube
 *                  (batch_stride 0)
 * batch_stride     elements between consecutive fields of GPH/dP (0 = shared cube)
 * geoid            (nlat, nlon) or NULL for 0; geoid_stride = {lat, lon} element strides
 * ring_tab         device, 9*nlat doubles built on the host from colatitude exactly as the
 *                  reference forms them: [0]=1-0.002644cos2t [1]=(1-0.0089cos2t)/6.245e6
 *                  [2]=N(t) [3]=N(t)(1-e^2) [4]=sin t [5]=cos t [6]=gs(t)
 *                  [7]=2(1.006803-0.06706cos^2 t) [8]=unused
 * lon_tab          device, 2*nlon doubles: cos(phi), sin(phi)
 * rad_e            datum(ELLIPSOID).rad_e [m]
 */
int mhb200_atmosphere_stokes(const mhb200_plan* plan, int dtype,
                             const void* GPH, const void* dP, int64_t batch_stride,
                             const double* field_scale,
                             const double* geoid, const int64_t geoid_stride[2],
                             const double* ring_tab, const double* lon_tab,
                             int nlev, int nbatch, int method, double rad_e,
                             const double* dfactor, const double* plm_table,
                             double* clm_out, double* slm_out,
                             void* workspace, size_t workspace_bytes, void* stream);

/*
 * Host-buffer form of mhb200_atmosphere_stokes for one field: GPH and dP are HOST cubes
 * (level, lat, lon) of type `dtype` (pinned memory gives full PCIe speed), clm/slm are HOST
 * outputs of (lmax+1)*(mmax+1) doubles.  The latitude axis is cut into `nbands` (<= 16) bands
 * and the host-to-device copy of band b+1 overlaps the kernel of band b.  The remaining
 * operands are as above (device tables).  Size the workspace with
 * mhb200_host_workspace_bytes(plan, nbands).  Returns after the results are on the host.
 */
size_t mhb200_host_workspace_bytes(const mhb200_plan* plan, int nbands);
int mhb200_atmosphere_stokes_host(const mhb200_plan* plan, int dtype,
                                  const void* GPH_host, const void* dP_host,
                                  const double* geoid, const int64_t geoid_stride[2],
                                  const double* ring_tab, const double* lon_tab,
                                  int nlev, int nbands, int method, double rad_e,
                                  const double* dfactor, const double* plm_table,
                                  double* clm_host, double* slm_host,
                                  void* workspace, size_t workspace_bytes, void* stream);

/*
 * The same with the geoid as a contiguous (nlat, nlon) HOST grid (what a driver holds): it is staged through a
 * plan-owned pinned buffer and copied behind the first band's cubes, so the staging of this pageable operand
 * overlaps the DMA of the cubes instead of preceding the call.
 */
int mhb200_atmosphere_stokes_host_geoid(const mhb200_plan* plan, int dtype,
                                        const void* GPH_host, const void* dP_host,
                                        const double* geoid_host,
                                        const double* ring_tab, const double* lon_tab,
                                        int nlev, int nbands, int method, double rad_e,
                                        const double* dfactor, const double* plm_table,
                                        double* clm_host, double* slm_host,
                                        void* workspace, size_t workspace_bytes, void* stream);

/*
 * gen_point_pressure.py:126-154,160-194.  Scattered points with disc areas:
 *   C_lm + i S_lm = dfactor[l] * sum_p Pdisc_l(alpha_p) (P/G)_p (R_p/rad_e)^(l+2) Plm(cos th_p) e^{i m phi_p}
 * phi, theta, area are (npts) device arrays; P is (nbatch, npts); G and R are addressed
 * with an element stride (0 broadcasts a 1-element operand).  dfactor (host) already holds
 * 4*pi*mmwe/(1+2l).
 * LMAX <= 61 runs one fused kernel (weights, e^{i m phi} and contraction in one launch) + the
 * fixed-order reduce; larger LMAX a preparation kernel + contraction + reduce.  Set
 * MHB200_POINTS_FUSED=0 to force the latter (tests compare the two).
 */
size_t mhb200_points_workspace_bytes(int64_t npts, int nbatch, int lmax, int mmax);
int mhb200_point_pressure_f64(int device, const double* P, const double* G, int64_t G_stride,
                              const double* R, int64_t R_stride,
                              const double* phi, const double* theta, const double* area,
                              int64_t npts, int nbatch, int lmax, int mmax,
                              double rad_e, const double* dfactor,
                              double* clm_out, double* slm_out,
                              void* workspace, size_t workspace_bytes, void* stream);

/*
 * Upstream producer of the atmosphere operator's inputs: hypsometric integration
 * over hybrid model levels, reanalysis/reanalysis_geopotential_heights.py:349-383.
 * T, Q            device float32 (nlev, ncol), bottom level first; ncol = nlat*nlon
 * SP, gh0         device float32 (ncol): surface pressure, surface geopotential (geopotential*GRAVITY)
 * A, B (nA), AI, BI (nAI)  host float64 hybrid coefficients at half levels / interfaces
 * divisor         outputs Z / divisor in float32 (9.80665 gives geopotential height, 1 leaves m^2/s^2)
 * Z_out, DIFF_out device float32 (nlev, ncol): geopotential at the levels, pressure difference
 * Arithmetic follows the reference's mixed float32/float64 rounding step by step.
 */
int mhb200_geopotential_levels_f32(int device, const float* T, const float* Q, const float* SP,
                                   const float* gh0, const double* A, const double* B, int nA,
                                   const double* AI, const double* BI, int nAI, int nlev, int64_t ncol,
                                   double R_dry, double Pmin, float divisor,
                                   float* Z_out, float* DIFF_out, void* stream);

/*
 * Probes of the library's pure host logic (no CUDA calls; used by the CPU tests).
 * mhb200_schedule_probe: the static equal-cost schedule for a problem shape.  seg receives
 *   5 ints per segment (field, degree block, order block, first unit, end unit), cta_begin the
 *   first segment of every CTA (ncta+1 entries); returns the segment count or -1.
 * mhb200_fold_probe: longitude symmetry pairing and the chunking chosen for a CTA shape (nth = 128
 *   or 256).  sets receives 4 ints per contraction column: the grid columns at phi, -phi, pi - phi,
 *   pi + phi (-1 where the fold level does not define them; repeated entries mean a self-paired
 *   column); *fold_out = 0, 1 or 2; returns the number of contraction columns or -1.
 */
int mhb200_schedule_probe(int nlat, int nchunks, int nlb, int nmb, int nbatch, int max_ctas, int mode,
                          int nlev, int fold, int h0, int h1, int* seg, int max_segs,
                          int* cta_begin, int max_ctas_buf, int* ncta_out);
int mhb200_fold_probe(int nlon, const double* phi, int nth, int* sets, int* unused, int* fold_out,
                      int* nchunks_out, int* kcf_out);

/*
 * mhb200_moment_schedule_probe: static schedule of the contracted-moment atmosphere kernel
 *   (csrc/stokes_moments.cu): the nb * nrings * nchunks (field, ring, 64-column chunk) units are split
 *   evenly over at most max_ctas CTAs; a slot is the part of one ring a CTA owns.  Fills
 *   cta_unit[ncta + 1], cta_slot[ncta], field_slot[nb + 1]; returns ncta (or -1), *nslots_out = slots.
 */
int mhb200_moment_schedule_probe(int nb, int nrings, int nchunks, int max_ctas, int* cta_unit,
                                 int* cta_slot, int* field_slot, int buf_ctas, int* nslots_out);

/*
 * mhb200_moment_layout_probe: chunk layout of the contracted-moment atmosphere kernel for nf contraction
 *   columns whose four fold sub-blocks are given in `sets` (4 ints per column, as mhb200_fold_probe returns
 *   them at fold level 2).  Fills col_grid[nchunks * 256] (grid column of every thread column, -1 = inactive:
 *   past the end, or a duplicate of a lower sub-block), box_start[nchunks * 4] (first longitude of every TMA
 *   box before alignment) and dir[4] (+1 ascending, -1 descending runs); returns 1 when every sub-block of
 *   every chunk is a contiguous run (the TMA feed applies), 0 when not, -1 on bad arguments.
 */
int mhb200_moment_layout_probe(int nf, const int* sets, int* col_grid, int* box_start, int* dir,
                               int max_chunks);

/* test hook: the fused point kernel's short sincos (Cody-Waite + Taylor, |a| < 5e5) on n device values */
int mhb200_sincos_probe(int device, const double* a, int64_t n, double* sin_out, double* cos_out, void* stream);

/* counters for bench.py: kernels launched by this library since load (all threads) */
uint64_t mhb200_launch_count(void);

/* measurement hook for bench.py: while enabled, every grid/point call brackets its dominant
 * kernel (atm_moment_kernel / stokes_ring_kernel / points_fused_kernel / ...) with CUDA events on the launching
 * stream; mhb200_profile_mean_ms() waits for them and returns the mean device time (ms) of the
 * launches recorded since enable (*n = how many), or a negative value if there were none */
void mhb200_profile_enable(int on);
double mhb200_profile_mean_ms(int* n);

#ifdef __cplusplus
}
#endif
#endif /* MHB200_H_ */
