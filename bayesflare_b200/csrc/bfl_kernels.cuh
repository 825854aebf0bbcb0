// bfl_kernels.cuh -- kernel launch interface shared by bfl_kernels.cu and bfl_abi.cu.
// sm_100a only; FP64 vector pipe (no tensor cores: a batched tiny-system solve + correlation).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bfl {

constexpr int MAXP = 6;      // polynomial terms (bgorder + 1) the general path is instantiated for
constexpr int MAXH = 8;      // hypotheses per plan
constexpr int MAXW = 255;    // window length limit

// one valid (unmasked) grid point of one hypothesis
struct TemplDesc {
  int32_t kind;
  int32_t reverse;
  double p[4];
};

// Device-side view of a plan.  All arrays are over the COMPACTED list of valid templates
// (grid points whose prior is finite), hypotheses stored back to back.
struct PlanView {
  int W, WP, P, G, nhyp;
  int hyp_begin[MAXH + 1];   // ranges in the compacted list
  int hyp_half[MAXH];        // lastHalfRange flag
  const double* raw;         // [G][WP] templates as the models define them
  const double* hat;         // [G][WP] templates orthogonalised against the background basis, unit norm
  const double* Kg;          // [G] -0.5*log(|m_perp|^2)
  const double* lp;          // [G] log prior (+ log amplitude prior)
  const double* lw;          // [G] log trapezium weight
  const int* gorig;          // [G] ravelled grid index inside its hypothesis
  const double* bg;          // [P][WP] background basis (monomials on [0,1])
  const double* bb;          // [P*(P+1)/2][WP] products b_p*b_q (upper triangle, row-major)
  const double* qb;          // [P][WP] orthonormalised background basis
  double logdetG;            // log det of the unit-variance background Gram matrix
  // ---- separable-grid form (see k_window_fast, SEP): template g = shape A[idxA[g]] + shape B[idxB[g]]
  // (orthogonalised, un-normalised), z_g = inorm[g] * (yA + yB).  Dense hypotheses have idxA = -1 and
  // their own normalised template as shape B.
  int sep_ok;                // the whole plan fits one CTA in this form
  int S;                     // number of shape vectors
  const double* shapes;      // [S][WP]
  const double* inorm;       // [G]
  const short* idxA;         // [G] index into the hypothesis' A shapes, or -1
  const short* idxB;         // [G] global shape index
  int hyp_A0[MAXH];          // global shape index of the hypothesis' first A shape
  int hyp_nA[MAXH];          // number of A shapes (0: dense hypothesis)
  // ---- the same decomposition of the RAW templates, for the weighted path (k_window_wsep): template g =
  // rshape A (rise) + rshape B (centre + decay) on disjoint supports, so the weighted norm separates too
  int wsep_ok;
  const double* rshapes;     // [S][WP]
  const short* sh_lo;        // [S] first tap of the shape's support (even)
  const short* sh_hi;        // [S] one past the last tap (<= W)
  const int* sh_g0;          // [S] templates [sh_g0, sh_g1) use shape s as their B shape (empty for A shapes)
  const int* sh_g1;
  int hyp_B0[MAXH];          // global shape index of the hypothesis' first B shape
  int hyp_nB[MAXH];          // number of B shapes
  // ---- split weighted form (bfl_wsplit.cu): the fixed vectors the streams 1/v and d/v are correlated with, packed by
  // (stream, support class) in that order; nullptr when the plan has no such form
  const double* wc_tab;      // packed vectors: whole-window rows of WP doubles, rise-half rows, decay-half rows
  const int* wc_off;         // [wc_nj] offset of a vector in wc_tab (doubles)
  const int* wc_row;         // [wc_nj] row of the sums array Yw it produces
  const int* wc_shrow;       // [S] first Yw row of shape s (f_0..f_{P-1}, e, r0), -1 for an empty shape
  int wc_n[2][3];            // vectors per (stream: 1/v, d/v) x (class: whole window, rise half, decay half)
  int wc_tab_len, wc_nj, wc_rows;   // rows: P(P+1)/2 of Q U Q^T, P of Q U d, P + 2 per non-empty shape
  // ---- per-plan shared-memory block of k_epilogue2 (tables, per-template weights, run lists), built once at
  // plan creation for noisepoly = 0 / 1; nullptr when the plan does not fit that kernel
  const void* epi2_block[2];
};

constexpr int SEP_MAXA = 16;   // A shapes per hypothesis the kernel keeps per thread in shared memory

struct SeriesView {
  const double* clc;    // dev [B][seglen]
  const double* cle;    // dev [B][seglen] or nullptr
  const double* sk;     // dev [B]
  int B;
  int64_t N, seg0, seglen, k0, k1;
  int const_var;
};

// per-launch scratch (device)
struct Scratch {
  double* sqrtu;   // [B]  sqrt(1/v)
  double* halflogv;// [B]  0.5*log(v)
  double* uconst;  // [B]  1/v
  double* dw;      // [B][seglen] d/v      (general path)
  double* uu;      // [B][seglen] v        (general path)
  double2* part;   // [nhyp][NC][B*nk] partial log-sum-exp (max, sum)
  double* bgabs;   // [B*nk] polynomial-only log Bayes factor (no sumlogsigma) or nullptr
  unsigned long long* tile_counters;  // [NC] dynamic tile scheduler state
  unsigned long long* split_counters; // [4] split form: self-resetting tile schedulers + exit counters
  double* Y;       // [S][B*nk] shape correlations (split form) or nullptr
  double* Yw;      // [wc_rows][Bc*nk] window sums of the split weighted form for a chunk of Bc curves, or nullptr
  int Yw_curves;   // Bc
  double* Yw_lnO;  // split weighted form: combine the hypotheses in k_wsolve and write the log odds here (or nullptr)
  int Yw_noisepoly;
  int NC;          // number of template chunks
  int gchunk;      // templates per chunk
  int KT;          // samples per thread of the fast kernel
  int sub;         // partial-sum slots per template chunk (1 for every current kernel)
};

struct LaunchStats { int64_t launches = 0; };

// one above-threshold region of one light curve (same layout as bfl_region in include/bfl.h)
struct RegionRec { int32_t curve, start, stop, argmax; double maxval; };

// arguments of the separable weighted kernels (k_window_wsep in bfl_kernels.cu, k_window_wsep2 in bfl_wsep2.cu)
struct WsepArgs {
  const double* dw; const double* uu;
  int64_t seglen, N, seg0, k0, k1;
  int B, tiles_per_curve;
  int64_t total_tiles;
  const double* rshapes; const short* sh_lo; const short* sh_hi; const int* sh_g0; const int* sh_g1;
  const double* qb; const double* lp; const double* lw; const short* idxA;
  int G, W, WP, S, nhyp, arows;
  int hyp_begin[MAXH + 1];
  int hyp_half[MAXH];
  int hyp_A0[MAXH], hyp_nA[MAXH], hyp_B0[MAXH], hyp_nB[MAXH];
  double logdetG;
  double2* part;
  double* bgabs;
};

// split weighted form (bfl_wsplit.cu): k_wcorr (window sums of both streams -> Yw) + k_wsolve (per-sample Cholesky,
// per-shape solves, per-template terms, log-sum-exp); Yw makes one round trip through HBM per chunk of curves
bool wsplit_window_supported(int W);
bool wsplit_fits(const PlanView& pv);
size_t wsplit_scratch_bytes(const PlanView& pv, int B, int64_t nk, int* chunk_curves);
// d_lnO not null: k_wsolve also combines the hypotheses (no partial sums, no k_combine) and writes the log odds
cudaError_t launch_window_wsplit(const PlanView& pv, const WsepArgs& a, double* Yw, int chunk_curves, int sms,
                                 double* d_lnO, int noisepoly, unsigned long long* counters, cudaStream_t s,
                                 LaunchStats* st);

cudaError_t launch_gen_templates(const TemplDesc* d_desc, int G, const double* d_mts, int W, int WP,
                                 double* d_raw, cudaStream_t s, LaunchStats* st);
cudaError_t launch_plan_orthonormalise(const double* d_raw, const double* d_qb, int G, int W, int WP, int P,
                                       double* d_hat, double* d_Kg, double* d_inorm, int normalise, cudaStream_t s,
                                       LaunchStats* st);
// A_i <- A_i - C for the separable form (rows [a0, a0+nA) minus row c of a [.][WP] array)
cudaError_t launch_subtract_row(double* d_rows, int a0, int nA, int c, int WP, cudaStream_t s, LaunchStats* st);

cudaError_t launch_prep(const SeriesView& sv, Scratch& sc, bool need_general, cudaStream_t s, LaunchStats* st);
// polynomial-only factor for interior constant-variance samples (fast path)
cudaError_t launch_bgproj(const PlanView& pv, const SeriesView& sv, Scratch& sc, cudaStream_t s, LaunchStats* st);
// fused window kernel, log-sum-exp mode (detector) -- interior, constant variance
cudaError_t launch_window_fast_lse(const PlanView& pv, const SeriesView& sv, Scratch& sc, cudaStream_t s,
                                   LaunchStats* st);
bool fast_can_fuse(const PlanView& pv, const Scratch& sc, int noisepoly);
cudaError_t launch_window_fast_fused(const PlanView& pv, const SeriesView& sv, Scratch& sc, int noisepoly,
                                     double* d_lnO, cudaStream_t s, LaunchStats* st);
// fused window kernel, full output mode for one hypothesis: out[gorig][B? no: single curve][nk]
cudaError_t launch_window_fast_full(const PlanView& pv, int ihyp, const SeriesView& sv, Scratch& sc,
                                    double sumlogsigma, double* d_out, cudaStream_t s, LaunchStats* st);
// general path (truncated windows at the series ends, or per-sample variance)
cudaError_t launch_window_general_lse(const PlanView& pv, const SeriesView& sv, Scratch& sc, bool edges_only,
                                      cudaStream_t s, LaunchStats* st);
cudaError_t launch_window_general_full(const PlanView& pv, int ihyp, const SeriesView& sv, Scratch& sc,
                                       bool edges_only, double sumlogsigma, double* d_out,
                                       cudaStream_t s, LaunchStats* st);
// weighted path: interior samples with per-sample variances, orthonormal-basis form
bool weighted_fits(const PlanView& pv, const Scratch& sc);
cudaError_t launch_window_weighted_lse(const PlanView& pv, const SeriesView& sv, Scratch& sc, cudaStream_t s,
                                       LaunchStats* st);
cudaError_t launch_window_weighted_full(const PlanView& pv, int ihyp, const SeriesView& sv, Scratch& sc,
                                        double sumlogsigma, double* d_out, cudaStream_t s, LaunchStats* st);
bool wsep_fits(const PlanView& pv);
cudaError_t launch_window_wsep_lse(const PlanView& pv, const SeriesView& sv, Scratch& sc, cudaStream_t s,
                                   LaunchStats* st);
// merge partial log-sum-exps, combine hypotheses into log odds
cudaError_t launch_combine(const PlanView& pv, const SeriesView& sv, const Scratch& sc, int noisepoly,
                           double* d_lnO, double* d_marg, cudaStream_t s, LaunchStats* st);

cudaError_t launch_whole_series(int P, const double* d_tm, const double* d_abg, const double* d_dt, const double* d_v,
                                const double* d_bgcross, const double* d_dbgr, const double* d_lp, int N, int NP, int G,
                                int half, double sls, double* d_out, double* d_out_bg, cudaStream_t s, LaunchStats* st);
cudaError_t launch_pe_grid(int kind, int reverse, const double* d_cts, const double* d_clc, int L,
                           const double* d_params, const double* d_logprior, int64_t G, int P,
                           const double* d_polyms, double sigma, int margbackground, double* d_post,
                           const double* d_sgcoef, int sgbins, cudaStream_t s, LaunchStats* st);

cudaError_t launch_logtrapz(const double* d_lx, const double* d_t, int64_t n, int64_t m, double* d_out,
                            cudaStream_t s, LaunchStats* st);

cudaError_t launch_threshold_edges(const double* d_lnO, int64_t n, double thresh, int64_t* d_starts,
                                   int64_t* d_stops, int* d_counts, int cap, cudaStream_t s, LaunchStats* st);
cudaError_t launch_region_max(const double* d_lnO, const int64_t* d_segs, int nsegs, double* d_maxval,
                              int64_t* d_maxidx, cudaStream_t s, LaunchStats* st);

cudaError_t launch_count_regions(const double* d_lnO, int B, int64_t n, double thresh, int64_t* d_counts,
                                 int64_t* d_totals, cudaStream_t s, LaunchStats* st);
cudaError_t launch_batch_regions(const double* d_lnO, int B, int64_t n, double thresh, RegionRec* d_recs, int cap,
                                 int64_t* d_counts, int64_t* d_totals, cudaStream_t s, LaunchStats* st);
cudaError_t launch_sg_resid(const double* d_y, int64_t N, int B, const double* d_m, int W, double* d_resid,
                            cudaStream_t s, LaunchStats* st);
cudaError_t launch_ps_sigma(const double2* d_X, int64_t N, int B, double estfrac, double* d_sk, cudaStream_t s,
                            LaunchStats* st);
cudaError_t launch_tv_sigma(const double* d_resid, int64_t N, int B, double sigma, int* d_counts, double* d_range,
                            double* d_sk, cudaStream_t s, LaunchStats* st);
cudaError_t launch_fp64_probe(double* d_sink, int iters, int blocks, cudaStream_t s, LaunchStats* st);

cudaError_t launch_sweep_simulate(unsigned long long seed, long long first, int B, int N, int H, double dt, double nstd,
                                  double sinusoid_amp, double snr_lo, double snr_hi, int inject, double* d_clc,
                                  double* d_truth, cudaStream_t s, LaunchStats* st);
cudaError_t launch_sweep_inject(int kind, int B, int N, double dt, const double* d_sk, double* d_truth, double* d_clc,
                                cudaStream_t s, LaunchStats* st);
cudaError_t launch_sweep_score(const double* d_lnO, int B, int64_t nk, int h, const double* d_truth, double thresh,
                               int nbins, double snr_max, int64_t* d_injected, int64_t* d_detected, int64_t* d_counters,
                               cudaStream_t s, LaunchStats* st);

bool fast_split_fits(const PlanView& pv);
bool split_window_supported(int W);
bool epi2_fits(const PlanView& pv);
size_t epi2_plan_block_bytes(const PlanView& pv);
cudaError_t launch_epilogue2_prepare(const PlanView& pv, int noisepoly, void* d_block, cudaStream_t s, LaunchStats* st);
cudaError_t launch_window_split(const PlanView& pv, const SeriesView& sv, Scratch& sc, int mode, int noisepoly,
                                double* d_lnO, cudaStream_t s, LaunchStats* st);
size_t fast_smem_bytes(int W, int gchunk, int KT);
size_t fast_sep_smem_bytes(int W, int G, int S, int KT, int arows);
bool fast_sep_fits(const PlanView& pv, int KT);
int pick_gchunk(int G, int W, int64_t total_samples, int sm_count, int* KT_out);

}  // namespace bfl
