/* bfl.h -- C ABI of libbfl.so: the B200 (sm_100a) sliding-window Bayesian odds-ratio engine.
 *
 * This is the drop-in boundary for BayesFlare's native layer.  The reference's native
 * boundary is Python -> Cython cpdef functions taking NumPy arrays
 * (bayesflare/stats/general.pyx:143-147, :249-252, :315, :576, :610) -> C
 * (bayesflare/stats/log_marg_amp_full.h:12-14).  Each entry point below names the
 * reference interface it replaces.  Plain pointers and sizes only; no torch / NumPy types.
 *
 * Conventions
 *   - every function returns an int status: 0 = ok, <0 = bfl_status; the message of the
 *     last failure on the calling thread is available from bfl_last_error();
 *   - nothing throws or aborts across the ABI;
 *   - "host" pointers are caller-owned C-contiguous float64 buffers (NumPy arrays);
 *     "dev" pointers are device allocations the caller owns (e.g. a torch tensor's
 *     data_ptr()) and work is enqueued on the context's stream without synchronising;
 *   - a context is bound to one device and one stream and is not thread-safe;
 *   - masked grid points (log prior = -inf) yield -inf exactly as the reference does.
 */
#ifndef BFL_H
#define BFL_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BFL_VERSION 100

typedef enum {
  BFL_OK = 0,
  BFL_ERR_ARG = -1,         /* bad argument (the reference prints and returns None) */
  BFL_ERR_CUDA = -2,        /* a CUDA runtime call failed; see bfl_last_error() */
  BFL_ERR_OOM = -3,         /* host or device allocation failed */
  BFL_ERR_UNSUPPORTED = -4, /* valid request outside what this build covers */
  BFL_ERR_NODEVICE = -5     /* no CUDA device: there is deliberately no CPU fallback */
} bfl_status;

/* template generators of bayesflare/models/model.py, evaluated on device */
typedef enum {
  BFL_MODEL_FLARE = 0,    /* Flare.model    model.py:197-252   params (t0, tauexp, taugauss, amp) */
  BFL_MODEL_EXPDECAY = 1, /* Expdecay.model model.py:560-607   params (t0, amp, tauexp, -)       */
  BFL_MODEL_IMPULSE = 2,  /* Impulse.model  model.py:694-736   params (t0, amp, -, -)            */
  BFL_MODEL_STEP = 3,     /* Step.model     model.py:943-982   params (t0, amp, -, -)            */
  BFL_MODEL_TRANSIT = 4,  /* Transit.model  model.py:387-436   params (t0, sigmag, tauf, amp)    */
  BFL_MODEL_GAUSSIAN = 5, /* Gaussian.model model.py:814-858   params (t0, sigma, amp, -)        */
  BFL_MODEL_TABLE = 6     /* caller-supplied templates (any other Model subclass)               */
} bfl_model_kind;

typedef struct bfl_ctx bfl_ctx;   /* device + stream + scratch memory */
typedef struct bfl_plan bfl_plan; /* a set of hypotheses on one window time base */

/* One hypothesis = one Model + its parameter grid (ravelled in C order, as
 * np.unravel_index does in bayes.py:348) + how its amplitude is marginalised. */
typedef struct {
  int32_t kind;          /* bfl_model_kind */
  int32_t reverse;       /* Model.reverse (Flare / Expdecay) */
  int32_t halfrange;     /* 1: amplitude in [0,inf) (lastHalfRange=1), 0: (-inf,inf) */
  int32_t ngrid;         /* number of grid points */
  const double* params;  /* host [ngrid][4], kind-specific order above; t0 = +inf means "window centre" */
  const double* logprior;/* host [ngrid]: Model.prior() + log amplitude prior (bayes.py:284-288,352-356); -inf masks */
  const double* logweight;/* host [ngrid]: log of the product trapezium weight of the grid point (general.pyx:576-607) */
  const double* table;   /* host [ngrid][W] templates when kind == BFL_MODEL_TABLE, else NULL */
} bfl_hypothesis;

/* ---- library / context -------------------------------------------------------------- */
int bfl_version(void);
const char* bfl_last_error(void);
int bfl_device_count(void);

/* device < 0 selects the current device.  stream == NULL creates a private stream;
 * otherwise `stream` is a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream). */
int bfl_ctx_create(int device, void* stream, bfl_ctx** ctx);
int bfl_ctx_destroy(bfl_ctx* ctx);
int bfl_ctx_sync(bfl_ctx* ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
int64_t bfl_ctx_launch_count(const bfl_ctx* ctx);
/* Measurement: when enabled, every launch of the fused window kernel is bracketed by CUDA
 * events on the context's stream.  bfl_ctx_window_time synchronises the stream, returns the
 * summed duration (ms) and the number of launches since the last call, and resets both.     */
int bfl_ctx_set_timing(bfl_ctx* ctx, int enable);
int bfl_ctx_window_time(bfl_ctx* ctx, double* total_ms, int64_t* launches);

/* ---- plan: templates + per-grid-point constants, built on device --------------------
 * Replaces the model/prior/cross-term set-up loop of Bayes.bayes_factors_marg_poly_bgd
 * (bayes.py:292-307 bgmodels, :347-356 templates, :367-393 model cross terms).
 *   W        bglen (odd), P = bgorder+1 (1..6)
 *   bgmodels host [P][W]  background basis, as the caller's NumPy computed it (bayes.py:297,307)
 *   mts      host [W]     time stamps the templates are evaluated on (bayes.py:261-276)     */
int bfl_plan_create(bfl_ctx* ctx, int W, int P, const double* bgmodels, const double* mts,
                    const bfl_hypothesis* hyps, int nhyp, bfl_plan** plan);
int bfl_plan_destroy(bfl_plan* plan);
/* device-generated templates of hypothesis ihyp -> host out[ngrid][W] (masked rows are zero).
 * Equivalent of Model.__call__(q, ts=mts).clc (model.py:133-153). */
int bfl_plan_templates(bfl_plan* plan, int ihyp, double* out);
/* number of correlation shapes the window kernel evaluates per sample for this plan: fewer than the
 * number of templates when a flare grid is held in separable form; -1 if the plan
 * is only available in dense form.  Used by bench.py to report the flops actually executed. */
int bfl_plan_shape_count(const bfl_plan* plan);

/* ---- Bayes.bayes_factors_marg_poly_bgd / _only (bayes.py:153-563) --------------------
 * replaces log_marg_amp_full_model / _2Dmodel / _background (general.pyx:143-365) and
 * log_marg_amp_full_C (log_marg_amp_full.c:3-68) for every sample and grid point.
 *   clc, cle host [N]  (cle may be NULL = zeros); sk noise sigma (bayes.py:239-245)
 *   sumlogsigma        sum_i log sqrt(sk^2+cle_i^2) as general.pyx:216-221 computes it
 *   out_lnB   host [ngrid][N] or NULL: lnBmargAmp of hypothesis ihyp (prior included)
 *   out_bg    host [N] or NULL: the polynomial-only factor (lnBmargBackground)           */
int bfl_window_lnB(bfl_ctx* ctx, bfl_plan* plan, int ihyp, const double* clc, const double* cle,
                   double sk, int64_t N, double sumlogsigma, double* out_lnB, double* out_bg);

/* ---- noise sigma on device (the pre-step of bayes.py:235-246) --------------------------
 * Savitzky-Golay detrend (noise.py:176-252, data.py:377-409) followed by either the
 * power-spectrum estimator (noise.py:10-44 with Lightcurve.psd, data.py:184-203; the FFT is
 * cuFFT) or the tail-veto estimator (noise.py:47-103: N-bin histogram CDF, np.unique, interp1d). */
typedef struct {
  int32_t method;       /* 0 = 'powerspectrum', 1 = 'tailveto' */
  int32_t bglen;        /* Savitzky-Golay window length; 0 = the curve is already detrended */
  double param;         /* psestfrac, or tvsigma */
  const double* sgcoef; /* host [bglen]: np.linalg.pinv(b)[0] as noise.py:241-245 computes it */
} bfl_noise_spec;

/* clc host [B][N] -> out_sk host [B] */
int bfl_noise_sigma(bfl_ctx* ctx, const bfl_noise_spec* spec, const double* clc, int32_t B, int64_t N,
                    double* out_sk);
/* device resident, asynchronous on the context's stream: d_clc dev [B][N] -> d_sk dev [B] */
int bfl_noise_sigma_dev(bfl_ctx* ctx, const bfl_noise_spec* spec, const double* d_clc, int32_t B, int64_t N,
                        double* d_sk);

/* ---- OddsRatioDetector.oddsratio (find.py:741-864), fused -----------------------------
 * hypothesis 0 of the plan is the signal; hypotheses 1.. are noise models; noisepoly adds
 * the polynomial-only noise model first, as find.py:774-782 does.  For a batch of B light
 * curves on the same time base:
 *   clc, cle host [B][N] (cle NULL = zeros)
 *   sk host [B] noise sigma per curve, or NULL to estimate it on device as `noise` describes
 *   out_sk host [B] or NULL: the sigmas used
 *   [k0, k1)               samples to produce: (0, N) for every sample, (W/2, N - W/2) when the
 *                          caller ignores the edges (find.py:846-851) -- the truncated-window
 *                          samples are then never computed
 *   out_lnO  host [B][k1-k0] log odds
 *   out_marg host [(nhyp+1)][B][k1-k0] or NULL: marginalised log Bayes factor of each hypothesis
 *                          (index nhyp = polynomial-only), without the sumlogsigma constant  */
int bfl_detector_odds(bfl_ctx* ctx, bfl_plan* plan, int noisepoly, const double* clc, const double* cle,
                      const double* sk, const bfl_noise_spec* noise, int32_t B, int64_t N, int64_t k0,
                      int64_t k1, double* out_lnO, double* out_marg, double* out_sk);

/* Same, device resident and asynchronous on the context's stream.  Supports time chunks:
 * the buffers hold samples [seg0, seg0+seglen) of series of true length N; log odds are
 * produced for samples [k0, k1) (k0-W/2 >= seg0 or seg0 == 0; likewise at the end) and
 * written to out_lnO[b*(k1-k0) + (k-k0)].
 *   d_clc, d_cle dev [B][seglen] (d_cle NULL = zeros; const_var = 1 promises cle is constant
 *   per curve so the translation-invariant fast path is taken), d_sk dev [B]               */
int bfl_detector_odds_dev(bfl_ctx* ctx, bfl_plan* plan, int noisepoly, const double* d_clc,
                          const double* d_cle, int const_var, const double* d_sk, int32_t B, int64_t N,
                          int64_t seg0, int64_t seglen, int64_t k0, int64_t k1, double* d_lnO);

/* Number of contiguous above-threshold regions (contiguous_regions, find.py:15-52) of each
 * curve of a device-resident batch d_lnO[B][n] -> d_counts[B] (int64, overwritten).
 * Asynchronous on the context's stream; feeds the multi-GPU detection-count reduction.      */
int bfl_count_regions_dev(bfl_ctx* ctx, const double* d_lnO, int32_t B, int64_t n, double thresh,
                          int64_t* d_counts);

/* Same, and in addition ACCUMULATES into d_totals[2] (device, int64; not cleared by the call): [0] the number
 * of regions in the batch, [1] the number of curves with at least one region -- the two counters a
 * multi-GPU search reduces over ranks (one small NCCL all-reduce per job instead of per batch). */
int bfl_count_regions_total_dev(bfl_ctx* ctx, const double* d_lnO, int32_t B, int64_t n, double thresh,
                                int64_t* d_counts, int64_t* d_totals);

/* ---- pipelined batch search ------------------------------------------------------------
 * The loop of scripts/kepler_analysis_script.py:340-455 (oddsratio + thresholder per star) for batches of
 * light curves on one time base, asynchronous: bfl_pipe_submit enqueues host->device copy, noise estimate,
 * all hypotheses, region extraction and the device->host copies of ONE batch on the stream of a pipeline
 * slot and returns; bfl_pipe_wait blocks until that batch's outputs are in host memory.  With depth >= 2
 * the copies of one batch overlap the kernels of its neighbours (host buffers must be page-locked for that;
 * pageable buffers work but serialise).  A slot is reused after `depth` submissions: its host outputs must
 * have been consumed by then.
 *   noise        how to estimate sigma on device when bfl_pipe_submit gets sk == NULL (copied; may be NULL)
 *   max_regions  capacity of the region list per batch
 * bfl_pipe_submit:
 *   clc  host [B][N]; sk host [B] or NULL
 *   out_lnO host [B][k1-k0] or NULL (NULL: the log odds never leave the device -- a search needs regions only)
 *   thresh NaN: no region extraction.  Otherwise contiguous_regions (find.py:15-52) of lnO > thresh per curve:
 *   out_regions host [max_regions] in arbitrary order (sort by curve, start), out_counts host [B + 3]:
 *   regions per curve, then {regions in the batch, curves with a region, records produced}
 *   out_sk host [B] or NULL: the sigmas used.   *job receives the job number for bfl_pipe_wait.           */
typedef struct bfl_pipe bfl_pipe;
typedef struct { int32_t curve, start, stop, argmax; double maxval; } bfl_region;
/* Region extraction alone, device resident and asynchronous on the context's stream: contiguous_regions of
 * d_lnO[B][n] > thresh -> d_regions[max_regions] (arbitrary order) and d_counts[B + 3] as described above.
 * Used by the multi-GPU paths, which gather region records -- not log odds -- over NCCL.                    */
int bfl_batch_regions_dev(bfl_ctx* ctx, const double* d_lnO, int32_t B, int64_t n, double thresh, bfl_region* d_regions,
                          int32_t max_regions, int64_t* d_counts);
int bfl_pipe_create(bfl_ctx* ctx, bfl_plan* plan, int noisepoly, const bfl_noise_spec* noise, int32_t B, int64_t N,
                    int64_t k0, int64_t k1, int32_t depth, int32_t max_regions, bfl_pipe** pipe);
int bfl_pipe_submit(bfl_pipe* pipe, const double* clc, const double* sk, double* out_lnO, double thresh,
                    bfl_region* out_regions, int64_t* out_counts, double* out_sk, int64_t* job);
int bfl_pipe_wait(bfl_pipe* pipe, int64_t job, int32_t* nregions);
int64_t bfl_pipe_launch_count(const bfl_pipe* pipe);
/* Jobs that reuse (slot, host buffers, threshold) are replayed as ONE CUDA graph from their second appearance on
 * (page-locked buffers only): number of graphs instantiated so far. */
int32_t bfl_pipe_graph_count(const bfl_pipe* pipe);
int bfl_pipe_destroy(bfl_pipe* pipe);

/* ---- injection sweep on device (scripts/flare_detection_efficiency.py:399-651; BASELINE config 5) ----
 * bfl_sweep_simulate_dev: B light curves of N <= 8192 samples, realisation indices first .. first+B-1 (the
 *   stream of a realisation depends on (seed, index) only -- counter-based Philox -- so a sweep sharded over
 *   ranks simulates the same curves whatever the number of ranks): white noise of standard deviation nstd,
 *   optional sinusoid (amplitude U[0, sinusoid_amp] nstd, period 2..31 d), one injected flare (script :321-335:
 *   tau_exp U[1800, 10800], tau_gauss U[0, 5400] redrawn until <= tau_exp, index randint(halfwin, N-halfwin-1),
 *   signal-to-noise U[snr_lo, snr_hi]) when inject == 1 -- scaled with nstd, median removed (data.py:271-277).
 *   inject == 2 / 3: the script's own recipe in two steps.  This call returns the BASE curve (noise + sinusoid, no
 *   median removal) and only draws and records the parameters of a flare (2) or a transit (3; script :291-312:
 *   tau_f U[1800, 14400], sigma_g U[0, 10800], redrawn until tau_f + 4 sigma_g <= 12 h); the caller estimates the
 *   noise sigma of the base curve as the script does (:429-434, bfl_noise_sigma_dev with the tail veto) and
 *   bfl_sweep_inject_dev adds the model rescaled to its signal-to-noise ratio with THAT sigma (:483-487).
 *   d_clc dev [B][N] out; d_truth dev [B][4] out: injection index (or -1), snr, then (tau_gauss, tau_exp) or
 *   (sigma_g, tau_f).
 * bfl_sweep_inject_dev: kind 1 = Flare.model, 2 = Transit.model at amplitude 1 on the whole series;
 *   d_clc[b] += model * snr / ((1 / d_sk[b]) sqrt(sum model^2)).  For a transit the injection index in d_truth is
 *   then set to -1: the script counts EVERY segment of the flare detector as a false positive there (:616-623).
 * bfl_sweep_score_dev: above-threshold runs of d_lnO[B][nk] (edge-trimmed by halfwin) expanded by +-2 samples
 *   and merged (script :521-580); the region holding the injection is a detection, every other one a false
 *   positive (:586-612).  ACCUMULATES (arrays not cleared): d_injected / d_detected [nbins] histograms over
 *   signal-to-noise in [0, snr_max] as np.histogram bins them, d_counters[2] = {false positives, regions}.
 * Both asynchronous on the context's stream.                                                            */
int bfl_sweep_simulate_dev(bfl_ctx* ctx, uint64_t seed, int64_t first, int32_t B, int32_t N, int32_t halfwin, double dt,
                           double nstd, double sinusoid_amp, double snr_lo, double snr_hi, int32_t inject, double* d_clc,
                           double* d_truth);
int bfl_sweep_inject_dev(bfl_ctx* ctx, int32_t kind, int32_t B, int32_t N, double dt, const double* d_sk,
                         double* d_truth, double* d_clc);
int bfl_sweep_score_dev(bfl_ctx* ctx, const double* d_lnO, int32_t B, int64_t nk, int32_t halfwin, const double* d_truth,
                        double thresh, int32_t nbins, double snr_max, int64_t* d_injected, int64_t* d_detected,
                        int64_t* d_counters);

/* ---- OddsRatioDetector.thresholder (find.py:915-1019) on device ----------------------
 * contiguous regions of lnO > thresh, expanded by `expand` and merged; per region the
 * maximum and its index.  lnO host [n]; segs host [max_segs][2] (start, stop), maxval /
 * maxidx host [max_segs]; *nsegs receives the number found (may exceed max_segs).       */
int bfl_threshold(bfl_ctx* ctx, const double* lnO, int64_t n, double thresh, int32_t expand,
                  int64_t* segs, double* maxval, int64_t* maxidx, int32_t max_segs, int32_t* nsegs);

/* ---- ParameterEstimationGrid.calculate_posterior (bayes.py:912-1055) -----------------
 * replaces log_likelihood_marg_background (general.pyx:610-656) ->
 * log_marg_amp_except_final_C (log_marg_amp_full.c:71-134) when margbackground != 0 and
 * log_likelihood_ratio (general.pyx:659-691) otherwise.  One light-curve chunk:
 *   cts, clc host [L]; grid: kind + params host [G][4] (+ logprior host [G]); P = bgorder+1
 *   polyms host [P][L]: linspace(0,1,L)**p as the caller's NumPy computed it (bayes.py:984-988)
 *   sgcoef host [sgbins] or NULL: when the chunk was Savitzky-Golay detrended, the filter taps (the same
 *   array bfl_noise_spec.sgcoef takes); every model gets m - savitzky_golay(m) as bayes.py:1018-1021 does
 *   out_post host [G] = log likelihood + log prior                                      */
int bfl_pe_grid(bfl_ctx* ctx, int32_t kind, int32_t reverse, const double* cts, const double* clc,
                int64_t L, const double* params, const double* logprior, int64_t G, int32_t P,
                const double* polyms, double sigma, int32_t margbackground, const double* sgcoef, int32_t sgbins,
                double* out_post);

/* ---- Bayes.bayes_factors_marg_poly_bgd with bglen = None, nsinusoids = 0 (bayes.py:247-437, the whole-series
 * branches :292-294, 330-331, 394-395, 408-409): ONE polynomial background over all N samples, the full-length
 * template slid along the series.  Replaces the same Cython/C chain as bfl_window_lnB.
 *   cts, clc, noisevar host [N] (noisevar = sk^2 + cle^2, bayes.py:312)
 *   params host [G][4], logprior host [G] (Model.prior + log amplitude prior; -inf rows stay -inf)
 *   bgmodels host [P][N] = linspace(0,1,N)**p as the caller's NumPy computed it (bayes.py:292-307)
 *   bgcross host [P][P] (upper triangle) = np.sum(bg_i bg_j / noisevar), dbgr host [P] = np.sum(clc / noisevar * bg_i):
 *   constants of the series, computed by the caller with NumPy exactly as the reference does
 *   out_lnB host [G][N] (may be NULL when G == 0)
 *   out_bg  host [1] or NULL: the polynomial-only factor of bayes_factors_marg_poly_bgd_only (bayes.py:439-563),
 *           one number: with a whole-series background it is the same for every sample                           */
int bfl_whole_series_lnB(bfl_ctx* ctx, int32_t kind, int32_t reverse, int32_t halfrange, const double* cts, const double* clc,
                         const double* noisevar, int64_t N, const double* params, const double* logprior, int64_t G,
                         int32_t P, const double* bgmodels, const double* bgcross, const double* dbgr, double sumlogsigma,
                         double* out_lnB, double* out_bg);

/* ---- logtrapz (general.pyx:576-607) over the leading axis of lx[n][m] -> out[m] ------ */
int bfl_logtrapz(bfl_ctx* ctx, const double* lx, const double* t, int64_t n, int64_t m, double* out);

/* ---- measurement helper: sustained FP64 FMA throughput (TFLOP/s) of this device,
 * the denominator of the window kernel's roofline (not in MEASURED_PEAKS.json).        */
int bfl_fp64_peak_probe(bfl_ctx* ctx, int32_t iters, double* tflops, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* BFL_H */
