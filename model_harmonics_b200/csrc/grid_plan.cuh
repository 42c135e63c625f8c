// Plan state shared by the grid kernels (stokes_grid.cu) and the contracted-moment path
// (stokes_moments.cu).  Internal to the library.
#pragma once
#include <cuda_runtime.h>
#include <mutex>
#include <vector>

namespace mhb {

constexpr int LBLK = 64;          // degrees per block
constexpr int MBLK = 64;          // orders per block
constexpr int MAX_PARTS = 16;

// work unit = (ring, longitude chunk), linearised as u = ring * nchunks + chunk
struct Segment {
  int t, lb, mb, u0, u1, pad;
};

struct MomentPlan;   // stokes_moments.cu

}  // namespace mhb

struct mhb200_plan;

namespace mhb {
// ---- contracted-moment path (stokes_moments.cu)
int moment_plan_create(mhb200_plan* pl, const double* phi);
void moment_plan_destroy(mhb200_plan* pl);
size_t moment_workspace_bytes(const mhb200_plan* pl, int nbatch, int launches);
bool moment_atm_eligible(const mhb200_plan* pl, int dtype, const void* GPH, const void* dP, long long batch_stride,
                         int nbatch, const double* plm_table);
int moment_atm_run(mhb200_plan* pl, int mode, int dtype, const void* GPH, const void* dP, long long batch_stride,
                   const double* field_scale_dev, const double* geoid, long long geo_s0, long long geo_s1,
                   const double* ring_tab, int nlev, int nbatch, double rad_e, const double* dfactor,
                   double* clm, double* slm, void* mom_ws, const int** flag_out, cudaStream_t st);
int moment_band_begin(mhb200_plan* pl, int nbands, int mode, const double* ring_tab, double rad_e, void* mom_ws,
                      const int** flag_out, cudaStream_t st);
int moment_band_launch(mhb200_plan* pl, int nbands, int mode, int dtype, const void* GPH, const void* dP,
                       const double* geoid, long long geo_s0, long long geo_s1, const double* ring_tab, int nlev,
                       double rad_e, int h0, int h1, int* slot_base, void* mom_ws, cudaStream_t st);
int moment_band_finish(mhb200_plan* pl, int nbands, int mode, int nslots, const double* dfactor, double* clm,
                       double* slm, void* mom_ws, cudaStream_t st);
bool moment_pressure_eligible(const mhb200_plan* pl, const double* plm_table);
int moment_pressure_run(mhb200_plan* pl, const double* P, const long long Ps[3], const double* G,
                        const long long Gs[3], const double* R, const long long Rs[3], int nbatch, double rad_e,
                        const double* dfactor, const double* dfactor_dev, double* clm, double* slm, void* mom_ws,
                        const int** flag_out, cudaStream_t st);
bool moment_layout(int nf, const std::vector<int>* const sets[4], int dir[4], std::vector<int>& col_grid,
                   std::vector<int>& box_start);
int moment_schedule(int nb, int nrings, int nchunks, int max_ctas, std::vector<int>& cta_unit,
                    std::vector<int>& cta_slot, std::vector<int>& field_slot, int* nslots);
}  // namespace mhb

using namespace mhb;

struct Schedule {
  int nbatch = -1, nlev = -1, mode = -1, h0 = -1, h1 = -1;
  int ncta = 0, nseg = 0;
  Segment* segs_dev = nullptr;
  int* cta_seg_begin_dev = nullptr;
  int* tile_slot_begin_dev = nullptr;
};

// longitude chunking + trig table for one CTA shape (variant 0: 256 threads, 1: 128 threads)
struct Variant {
  int nth = 0, ctas_per_sm = 1;
  int kc = 0, kcf = 0, nchunks = 0, nks_total = 0;   // thread columns / contraction columns per chunk
  double* trig = nullptr;
  int* phys = nullptr;                               // [nchunks][kc] grid column of each thread column
  int *pf_cols = nullptr, *pf_count = nullptr;       // L2 prefetch line lists (GridParams::pf_cols)
  Schedule sched;
};

struct mhb200_plan {
  int device, nlat, nlon, lmax, mmax;
  int nlb, nmb, Mpad, ntiles, num_sms;
  Variant var[2];
  double *f1 = nullptr, *f2 = nullptr, *pmm = nullptr, *sq = nullptr, *rescale = nullptr, *x = nullptr,
         *w = nullptr;
  double* ckpt = nullptr;  // Legendre restart states (nlb > 1)
  double* bmat = nullptr;  // binomial-moment expansion matrices (nlb == 1): [BC05 | SW02][4][8][32] B fragments
  // longitude mirror folding: contraction column i pairs grid column colp[i] with its mirror
  // colq[i] (phi_q = -phi_p mod 2 pi; colq == colp for a self-mirrored column)
  int fold = 0;            // 0: none, 1: phi -> -phi, 2: also phi -> pi - phi
  std::vector<int> colp, colq, colp2, colq2;
  // host-pipelined atmosphere path: per-band schedules, device cubes, copy stream
  Schedule band_sched[MAX_PARTS];
  void* cube_dev[2] = {nullptr, nullptr};
  size_t cube_bytes = 0;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t band_ev[MAX_PARTS] = {};
  cudaEvent_t done_ev = nullptr;
  double* out_dev = nullptr;      // 2 x (lmax+1)(mmax+1)
  double* out_pinned = nullptr;
  double *geo_dev = nullptr, *geo_pinned = nullptr;   // staging of a host geoid (mhb200_atmosphere_stokes_host_geoid)
  int* lb_first_tile_dev = nullptr;
  std::vector<int> lb_first_tile;
  std::mutex mu;
  mhb::MomentPlan* mom = nullptr;   // contracted-moment path (null: grid not eligible)
};

