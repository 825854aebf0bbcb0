// bfl_kernels.cu -- hand-written sm_100a kernels for the sliding-window Bayesian odds ratio.
//
// What is computed (reference: bayesflare/stats/bayes.py:153-563, general.pyx:143-365,
// log_marg_amp_full.c:3-134, finder/find.py:741-864): for every time sample k and every
// model grid point g, the logarithm of the amplitude-marginalised likelihood ratio of
// "template g + polynomial background" against Gaussian noise on a W-sample window
// centred on k; then a trapezium log-sum-exp over g and the combination of hypotheses.
//
// Three paths, all fp64 on the vector pipe:
//
//  * FAST path (interior samples, noise variance constant along the curve).  The Gram
//    matrix of [polynomials | template] is translation invariant, so its factorisation is
//    done once per plan: the template is projected off an orthonormalised polynomial basis
//    and normalised (m_hat).  Per (k,g) the whole marginalisation collapses to ONE
//    correlation z = sqrt(u) * sum_w m_hat[w] d[k-h+w] followed by
//        lnB_g(k) - lnB_poly(k) = 0.5 log v - 0.5 log|m_perp|^2 + { 0.5 log(pi/2) + z^2/2 + log erfc(-z/sqrt2)   (half range)
//                                                                  { 0.5 log(2 pi) + z^2/2                         (full range)
//    which is algebraically identical to the reference's sequential elimination
//    (log_marg_amp_full.c:29-66) but better conditioned (no Hilbert-like pivots).
//    The thread keeps its KT+W-1 data window in registers; templates are broadcast from
//    shared memory; the log-sum-exp over the grid is online, so per-grid values never
//    reach HBM in detector mode.  A flare grid is separable (rise + decay shapes), and large
//    batches run the split form: k_corr_shapes (shape correlations) + k_epilogue (grid sum in the
//    linear domain, hypotheses combined).
//
//  * WEIGHTED path (interior samples of curves with per-sample variance, i.e. real light curves
//    with flux errors): the same integral with a per-sample P x P Cholesky in the orthonormal
//    basis.  Window lengths 21 / 33 / 55 and the default grid: the split form of bfl_wsplit.cu (k_wcorr + k_wsolve);
//    otherwise k_window_wsep (separable, one kernel) or k_window_weighted (dense / full output).
//
//  * GENERAL path (first/last W/2 samples where the reference truncates the window,
//    bayes.py:320-327).  These systems are ill conditioned and the reference's answer depends on
//    its rounding: every cross term is rebuilt per sample in the reference's own summation order
//    (NumPy pairwise) and the elimination is replayed in the reference's order without FMA
//    contraction.
//
// Further down: plan kernels (templates, orthonormalisation), noise estimate (Savitzky-Golay,
// periodogram, tail veto), parameter-estimation grid, thresholding, the injection sweep.
#include "bfl_kernels.cuh"

#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include <type_traits>

#include "bfl_devmath.cuh"

namespace bfl {


// ======================================================================================
// plan kernels
// ======================================================================================
// Templates on device: models/model.py.  One thread per (template, window sample).
__global__ void k_gen_templates(const TemplDesc* __restrict__ desc, int G, const double* __restrict__ mts,
                                int W, int WP, double* __restrict__ raw) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= G * WP) return;
  const int g = idx / WP, w = idx - g * WP;
  if (w >= W) { raw[idx] = 0.0; return; }
  const TemplDesc d = desc[g];
  if (d.kind == 6) return;  // TABLE: rows uploaded by the host
  const double ts = mts[w];
  double t0 = d.p[0];
  const bool centre = isinf(t0) && t0 > 0;
  if (centre) t0 = mts[W / 2];  // ts[int(len(ts)/2.)], model.py:222-223
  double f = 0.0;
  switch (d.kind) {
    case 0: {  // Flare (t0, tauexp, taugauss, amp), model.py:197-252
      const double tauexp = d.p[1], taugauss = d.p[2], amp = d.p[3];
      const bool before = ts < t0, after = ts > t0;
      if (ts == t0) f = amp;
      if (taugauss > 0 && (d.reverse ? after : before)) {
        const double dt = ts - t0;
        f = amp * exp(-(dt * dt) / (2.0 * (taugauss * taugauss)));
      }
      if (tauexp > 0) {
        if (d.reverse && before) f = amp * exp((ts - t0) / tauexp);
        if (!d.reverse && after) f = amp * exp(-(ts - t0) / tauexp);
      }
    } break;
    case 1: {  // Expdecay (t0, amp, tauexp), model.py:560-607
      const double amp = d.p[1], tauexp = d.p[2];
      if (ts == t0) f = amp;
      if (tauexp > 0) {
        if (d.reverse && ts < t0) f = amp * exp((ts - t0) / tauexp);
        if (!d.reverse && ts > t0) f = amp * exp(-(ts - t0) / tauexp);
      }
    } break;
    case 2: {  // Impulse (t0, amp): finite t0 is an offset from ts[0]; nearest sample, model.py:694-736
      const double amp = d.p[1];
      if (!centre) t0 = mts[0] + d.p[0];
      int best = 0;
      double bd = fabs(mts[0] - t0);
      for (int i = 1; i < W; i++) {
        const double a = fabs(mts[i] - t0);
        if (a < bd) { bd = a; best = i; }
      }
      f = (w == best) ? amp : 0.0;
    } break;
    case 3: {  // Step (t0, amp), model.py:943-982
      f = (ts > t0) ? d.p[1] : 0.0;
    } break;
    case 4: {  // Transit (t0, sigmag, tauf, amp), model.py:387-436
      const double sigmag = d.p[1], tauf = d.p[2], amp = d.p[3];
      f = -1.0 * amp;
      if (ts < t0 - tauf) {
        const double x = ts - (t0 - tauf);
        f = sigmag > 0 ? (-1.0 * amp) * exp(-(x * x) / (2.0 * (sigmag * sigmag))) : 0.0;
      }
      if (ts > t0 + tauf) {
        const double x = ts - (t0 + tauf);
        f = sigmag > 0 ? (-1.0 * amp) * exp(-(x * x) / (2.0 * (sigmag * sigmag))) : 0.0;
      }
    } break;
    case 5: {  // Gaussian (t0, sigma, amp), model.py:814-858
      const double sigma = d.p[1], amp = d.p[2];
      if (sigma == 0) {
        // f[np.amin(tm0) == tm0] = amp : marks the minimum of ts - t0 (the earliest sample)
        double mn = mts[0] - t0;
        for (int i = 1; i < W; i++) mn = fmin(mn, mts[i] - t0);
        f = ((ts - t0) == mn) ? amp : 0.0;
      } else {
        const double dt = ts - t0;
        f = amp * exp(-(dt * dt) / (2.0 * (sigma * sigma)));
      }
    } break;
    default: break;
  }
  raw[idx] = f;
}

// One thread per template: remove the projection on the orthonormal background basis
// (twice, for orthogonality at rounding level); optionally normalise; record
// -0.5 log|m_perp|^2 and 1/|m_perp|.
__global__ void k_plan_orthonormalise(const double* __restrict__ raw, const double* __restrict__ qb, int G,
                                      int W, int WP, int P, double* __restrict__ hat, double* __restrict__ Kg,
                                      double* __restrict__ inorm, int normalise) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  double* h = hat + (size_t)g * WP;
  const double* r = raw + (size_t)g * WP;
  for (int w = 0; w < WP; w++) h[w] = r[w];
  for (int pass = 0; pass < 2; pass++) {
    for (int j = 0; j < P; j++) {
      const double* q = qb + (size_t)j * WP;
      double c = 0.0;
      for (int w = 0; w < W; w++) c = fma(q[w], h[w], c);
      for (int w = 0; w < W; w++) h[w] = fma(-c, q[w], h[w]);
    }
  }
  double X = 0.0;
  for (int w = 0; w < W; w++) X = fma(h[w], h[w], X);
  const double inv = 1.0 / sqrt(X);
  if (normalise)
    for (int w = 0; w < W; w++) h[w] *= inv;
  if (Kg) Kg[g] = -0.5 * log(X);
  if (inorm) inorm[g] = inv;
}

__global__ void k_subtract_row(double* rows, int a0, int nA, int c, int WP) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nA * WP) return;
  const int r = i / WP, w = i - r * WP;
  rows[(size_t)(a0 + r) * WP + w] -= rows[(size_t)c * WP + w];
}

cudaError_t launch_gen_templates(const TemplDesc* d_desc, int G, const double* d_mts, int W, int WP,
                                 double* d_raw, cudaStream_t s, LaunchStats* st) {
  if (G == 0) return cudaSuccess;
  const int n = G * WP;
  k_gen_templates<<<(n + 255) / 256, 256, 0, s>>>(d_desc, G, d_mts, W, WP, d_raw);
  st->launches++;
  return cudaGetLastError();
}

cudaError_t launch_plan_orthonormalise(const double* d_raw, const double* d_qb, int G, int W, int WP, int P,
                                       double* d_hat, double* d_Kg, double* d_inorm, int normalise, cudaStream_t s,
                                       LaunchStats* st) {
  if (G == 0) return cudaSuccess;
  k_plan_orthonormalise<<<(G + 63) / 64, 64, 0, s>>>(d_raw, d_qb, G, W, WP, P, d_hat, d_Kg, d_inorm, normalise);
  st->launches++;
  return cudaGetLastError();
}

cudaError_t launch_subtract_row(double* d_rows, int a0, int nA, int c, int WP, cudaStream_t s, LaunchStats* st) {
  if (nA <= 0) return cudaSuccess;
  k_subtract_row<<<(nA * WP + 127) / 128, 128, 0, s>>>(d_rows, a0, nA, c, WP);
  st->launches++;
  return cudaGetLastError();
}

// ======================================================================================
// per-launch preparation: per-curve scalars and (general path) whitened series
// ======================================================================================
__global__ void k_prep_scalars(const double* __restrict__ sk, const double* __restrict__ cle, int64_t seglen,
                               int B, double* sqrtu, double* halflogv, double* uconst,
                               unsigned long long* tile_counters, int ncounters) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < ncounters) tile_counters[b] = 0ULL;   // dynamic tile scheduler of the window kernel
  if (b >= B) return;
  const double s = sk[b];
  const double e = cle ? cle[(size_t)b * seglen] : 0.0;
  const double v = __dadd_rn(__dmul_rn(s, s), __dmul_rn(e, e));   // bayes.py:312, no FMA contraction
  const double u = 1.0 / v;
  uconst[b] = u;
  sqrtu[b] = sqrt(u);
  halflogv[b] = 0.5 * log(v);
}

__global__ void k_prep_whiten(const double* __restrict__ clc, const double* __restrict__ cle,
                              const double* __restrict__ sk, int64_t seglen, int B, double* __restrict__ dw,
                              double* __restrict__ uu) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= seglen * B) return;
  const int b = (int)(i / seglen);
  const double s = sk[b];
  const double e = cle ? cle[i] : 0.0;
  const double v = __dadd_rn(__dmul_rn(s, s), __dmul_rn(e, e));   // sk**2 + cle**2 as NumPy rounds it
  dw[i] = __ddiv_rn(clc[i], v);   // bayes.py:403
  uu[i] = v;   // the variance itself: the general path divides, as bayes.py:321-327 does
}

cudaError_t launch_prep(const SeriesView& sv, Scratch& sc, bool need_general, cudaStream_t s, LaunchStats* st) {
  k_prep_scalars<<<(std::max(sv.B, sc.NC) + 127) / 128, 128, 0, s>>>(sv.sk, sv.cle, sv.seglen, sv.B, sc.sqrtu, sc.halflogv,
                                                                   sc.uconst, sc.tile_counters, sc.NC);
  st->launches++;
  if (need_general) {
    const int64_t n = sv.seglen * sv.B;
    k_prep_whiten<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(sv.clc, sv.cle, sv.sk, sv.seglen, sv.B, sc.dw, sc.uu);
    st->launches++;
  }
  return cudaGetLastError();
}

// SM count of the CURRENT device (a process may drive several devices of different sizes)
static int sm_count_current() {
  static int cache[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) { int n = 0; cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); return n; }
  if (!cache[dev]) cudaDeviceGetAttribute(&cache[dev], cudaDevAttrMultiProcessorCount, dev);
  return cache[dev];
}

// ======================================================================================
// FAST path
// ======================================================================================
struct FastArgs {
  const double* data;     // [B][seglen]
  const double* sqrtu;    // [B]
  const double* halflogv; // [B]
  int64_t seglen, N, seg0, k0, k1;
  int B, tiles_per_curve;   // tiles_per_curve: warp tiles (32*KT samples) per light curve
  int64_t total_wtiles;
  unsigned long long* tile_counters;   // [NC] next warp tile to hand out, per template chunk
  const double* hat;      // [G][WP]
  const double* Kg;       // [G]
  const double* lp;       // [G]
  const double* lw;       // [G]
  const int* gorig;       // [G]
  int G, W, WP, gchunk, NC, nhyp;
  int hyp_begin[MAXH + 1];
  int hyp_half[MAXH];
  double2* part;          // LSE mode
  double* out_full;       // FULL mode: [ngrid][nk]
  const double* bgabs;    // FULL mode: [nk]
  double sls;             // FULL mode: sumlogsigma
  int ihyp;               // FULL mode
  double* lnO;            // FUSED mode (MODE 2): log odds written directly, [B][nk]
  int noisepoly;          // FUSED mode
  // separable-grid form (SEP kernels)
  const double* shapes; const double* inorm; const short* idxA; const short* idxB;
  int S, arows;
  int hyp_A0[MAXH];
  int hyp_nA[MAXH];
};

constexpr int FAST_TPB = 128;

// z for KT consecutive samples against one template (register window, compile-time W).
// Two partial sums per sample, pairs of taps alternate between them: the order is the same for
// every KT, so results do not depend on the tiling.
template <int WT, int KT>
__device__ __forceinline__ void correlate_reg(const double* __restrict__ t, const double* win, double (&z)[KT]) {
  constexpr int NPAIR = WT / 2;     // WT is odd: NPAIR pairs + one last tap
  double acc[KT][2];
#pragma unroll
  for (int j = 0; j < KT; j++) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
  const double2* t2 = reinterpret_cast<const double2*>(t);
#pragma unroll
  for (int w2 = 0; w2 < NPAIR; w2++) {
    const double2 m = t2[w2];
#pragma unroll
    for (int j = 0; j < KT; j++) {
      acc[j][w2 & 1] = fma(m.x, win[2 * w2 + j], acc[j][w2 & 1]);
      acc[j][w2 & 1] = fma(m.y, win[2 * w2 + 1 + j], acc[j][w2 & 1]);
    }
  }
  const double ml = t[WT - 1];
#pragma unroll
  for (int j = 0; j < KT; j++) z[j] = fma(ml, win[WT - 1 + j], acc[j][NPAIR & 1]) + acc[j][(NPAIR + 1) & 1];
}

// WT > 0: compile-time window length with the data window in registers;
// WT == 0: run-time window length (any odd W <= MAXW), data window read from shared memory, KT = 1.
#ifndef BFL_MINBLOCKS
#define BFL_MINBLOCKS 2
#endif
// Persistent kernel: the grid is sized to the number of CTAs the chip can hold; each CTA stages
// its template chunk and the function tables into shared memory ONCE, then its warps walk over
// "warp tiles" (32*KT consecutive samples of one light curve) independently of each other --
// each warp stages its own data window (+ halo) into a private shared-memory slice, so the tile
// loop contains no block-level barrier.
// SEP: the plan's separable-grid form (PlanView::shapes): per hypothesis the A-shape correlations are
// computed once per tile into a per-thread shared-memory slice, then every template costs one B-shape
// correlation per distinct B (amortised over the templates that share it) plus z = inorm * (yA + yB).
template <int WT, int KT, int MODE, bool SEP>
__global__ void __launch_bounds__(FAST_TPB, BFL_MINBLOCKS) k_window_fast(const FastArgs a) {
  extern __shared__ double smem[];
  const int W = (WT > 0) ? WT : a.W;
  const int H = W / 2;
  const int WP = a.WP;
  constexpr int TKW = 32 * KT;                   // samples per warp tile
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunk = blockIdx.y;
  const int gb = chunk * a.gchunk;
  const int ge = min(a.G, gb + a.gchunk);
  const int wstride = (TKW + 2 * H + 2) & ~1;    // doubles per warp slice

  const int nrows = SEP ? a.S : a.gchunk;                // shape rows (SEP) or template rows staged
  double* sdata = smem;                                  // [WPB][wstride]
  double* stmpl = smem + (FAST_TPB / 32) * wstride;      // [nrows + 1][WP] (last row: zeros)
  double* sK = stmpl + (size_t)(nrows + 1) * WP;         // [gchunk]  Kg + lp (+ lw)
  double* sphi = sK + ((a.gchunk + 1) & ~1);             // [(DEG+1) * NI] phi(z) table
  double* st64 = sphi + BFL_PHI_CM_SIZE;                 // [64] 2^(j/64)
  double* sInorm = st64 + 64;                            // SEP: [gchunk] 1/|m_perp|
  short* sIdx = reinterpret_cast<short*>(sInorm + (SEP ? ((a.gchunk + 1) & ~1) : 0));   // SEP: [gchunk][2] (A, B)
  double* ysm = reinterpret_cast<double*>(sIdx + (SEP ? ((2 * a.gchunk + 7) & ~7) : 0)); // SEP: [SEP_MAXA][KT][FAST_TPB]

  // ---- once per CTA: templates of this chunk (16-byte copies) and the function tables ----
  {
    const double2* src = reinterpret_cast<const double2*>(SEP ? a.shapes : a.hat + (size_t)gb * WP);
    double2* dst = reinterpret_cast<double2*>(stmpl);
    const int nr = SEP ? a.S : (ge - gb);
    const int n2 = nr * WP / 2;
    for (int i = tid; i < n2; i += FAST_TPB) dst[i] = __ldg(src + i);
    for (int i = tid; i < WP; i += FAST_TPB) stmpl[(size_t)nr * WP + i] = 0.0;   // zero row
    for (int i = tid; i < ge - gb; i += FAST_TPB) {
      double k = a.Kg[gb + i] + a.lp[gb + i];
      if (MODE != 1) k += a.lw[gb + i];
      sK[i] = k;
      if (SEP) {
        sInorm[i] = a.inorm[gb + i];
        sIdx[2 * i] = a.idxA[gb + i];
        sIdx[2 * i + 1] = a.idxB[gb + i];
      }
    }
    if (tid < 64) st64[tid] = exp2((double)tid * (1.0 / 64.0));
    for (int i = tid; i < (BFL_PHI_DEG + 1) * BFL_PHI_NI; i += FAST_TPB) {   // [coef][interval] copy
      const int cidx = i / BFL_PHI_NI, kk = i - cidx * BFL_PHI_NI;
      sphi[i] = BFL_PHI_TAB[kk * BFL_PHI_ROW + cidx];
    }
  }
  __syncthreads();

  // templates of this chunk, restricted to one hypothesis in FULL mode
  int g0 = gb, g1 = ge;
  if (MODE == 1) { g0 = max(gb, a.hyp_begin[a.ihyp]); g1 = min(ge, a.hyp_begin[a.ihyp + 1]); }
  if (g0 >= g1) return;
  int h0 = 0;
  while (a.hyp_begin[h0 + 1] <= g0) h0++;
  const int64_t nk = a.k1 - a.k0;
  double* swin = sdata + warp * wstride;

  // dynamic scheduling: every warp takes the next warp tile from a global counter, so the
  // tail is one tile long however the tile count divides among the resident warps
  for (;;) {
  long long wt = 0;
  if (lane == 0) wt = (long long)atomicAdd(a.tile_counters + chunk, 1ULL);
  wt = __shfl_sync(0xffffffffu, wt, 0);
  if (wt >= a.total_wtiles) break;
  const int tile = (int)(wt % a.tiles_per_curve);
  const int b = (int)(wt / a.tiles_per_curve);
  const int64_t kbase = a.k0 + (int64_t)tile * TKW;

  // ---- stage this warp's light-curve slice + halo (coalesced), pre-scaled by sqrt(1/v) ----
  {
    const double su = a.sqrtu[b];
    const double* d = a.data + (size_t)b * a.seglen;
    const int64_t lo = a.seg0, hi = min(a.N, a.seg0 + a.seglen);
    __syncwarp();
    for (int i = lane; i < TKW + 2 * H; i += 32) {
      const int64_t sidx = kbase - H + i;
      double v = 0.0;
      if (sidx >= lo && sidx < hi) v = __ldg(d + (sidx - a.seg0)) * su;
      swin[i] = v;
    }
    __syncwarp();
  }

  // ---- the thread's data window ----
  double win[(WT > 0) ? (WT + KT - 1) : 1];
  if constexpr (WT > 0) {
#pragma unroll
    for (int e = 0; e < WT + KT - 1; e++) win[e] = swin[lane * KT + e];
  }

  const double hlv = a.halflogv[b];
  const int64_t kfirst = kbase + (int64_t)lane * KT;

  bool valid[KT];
#pragma unroll
  for (int j = 0; j < KT; j++) {
    const int64_t k = kfirst + j;
    valid[j] = (k < a.k1) && (k >= H) && (k < a.N - H);   // interior samples only
  }

  auto correlate = [&](int gl, double (&z)[KT]) {
    const double* t = stmpl + (size_t)gl * WP;
    if constexpr (WT > 0) {
      correlate_reg<WT, KT>(t, win, z);
    } else {
      double a0 = 0.0, a1 = 0.0;
      const double* dd = swin + lane;
      int w = 0;
      for (; w + 1 < W; w += 2) { a0 = fma(t[w], dd[w], a0); a1 = fma(t[w + 1], dd[w + 1], a1); }
      if (w < W) a0 = fma(t[w], dd[w], a0);
      z[0] = a0 + a1;
    }
  };

  int h = h0;
  double mx[KT], sm[KT];
#pragma unroll
  for (int j = 0; j < KT; j++) { mx[j] = LSE_EMPTY; sm[j] = 0.0; }

  // MODE 2 (fused): the hypotheses are combined in registers, find.py:853-862 -- numerator = the
  // signal's marginal, denominator = sum over [polynomial-only (= 1 in these units), noise models].
  // Both are kept as (max, sum) pairs and merged with one cheap exp per hypothesis; a single log per
  // sample is taken at the end.  Needs every template of the plan in this CTA's chunk (NC == 1).
  double n_m[KT], n_s[KT], d_m[KT], d_s[KT];
#pragma unroll
  for (int j = 0; j < KT; j++) {
    n_m[j] = LSE_EMPTY; n_s[j] = 0.0;
    d_m[j] = a.noisepoly ? 0.0 : LSE_EMPTY; d_s[j] = a.noisepoly ? 1.0 : 0.0;
  }
  auto flush = [&](int hh) {
    const double hc = (a.hyp_half[hh] ? 0.5 * (BFL_LNPI - BFL_LN2) : 0.5 * (BFL_LN2 + BFL_LNPI)) + hlv;
    if (MODE == 0) {
      double2* p = a.part + ((size_t)(hh * a.NC + chunk) * a.B + b) * nk;
#pragma unroll
      for (int j = 0; j < KT; j++) {
        if (valid[j]) p[kfirst + j - a.k0] = make_double2(mx[j] + hc, sm[j]);
        mx[j] = LSE_EMPTY; sm[j] = 0.0;
      }
    } else if (MODE == 2) {
#pragma unroll
      for (int j = 0; j < KT; j++) {
        const double x = mx[j] + hc;
        if (hh == 0) { n_m[j] = x; n_s[j] = sm[j]; }
        else {
          const double d = x - d_m[j];
          const double e = exp_nonpos(-fabs(d), st64);
          const bool up = d > 0.0;
          d_s[j] = up ? fma(d_s[j], e, sm[j]) : fma(sm[j], e, d_s[j]);
          d_m[j] = up ? x : d_m[j];
        }
        mx[j] = LSE_EMPTY; sm[j] = 0.0;
      }
    }
  };

  // hypotheses -> (SEP: passes over groups of A shapes) -> templates
  double yB[KT];
  double zc[KT];
  int h_last = h0;
  while (a.hyp_begin[h_last + 1] < g1) h_last++;
  for (h = h0; h <= h_last; h++) {
    const int glo = max(g0, a.hyp_begin[h]), ghi = min(g1, a.hyp_begin[h + 1]);
    if (glo >= ghi) continue;
    const int nA = SEP ? a.hyp_nA[h] : 0;
    const int npass = (nA + SEP_MAXA - 1) / SEP_MAXA > 1 ? (nA + SEP_MAXA - 1) / SEP_MAXA : 1;
    for (int pass = 0; pass < npass; pass++) {
      const int ia0 = pass * SEP_MAXA, ia1 = min(nA, ia0 + SEP_MAXA);
      if constexpr (SEP) {
        // correlations with this group's A shapes, kept per thread in shared memory
        for (int ia = ia0; ia < ia1; ia++) {
          double y[KT];
          correlate(a.hyp_A0[h] + ia, y);
#pragma unroll
          for (int j = 0; j < KT; j++) ysm[((ia - ia0) * KT + j) * FAST_TPB + tid] = y[j];
        }
      }
      int curB = -1;
      for (int g = glo; g < ghi; g++) {
        if constexpr (SEP) {
          const int ia = sIdx[2 * (g - gb)], ib = sIdx[2 * (g - gb) + 1];
          if (nA > 0 && (ia < ia0 || ia >= ia1)) continue;         // another pass owns this template
          if (ib != curB) { curB = ib; correlate(ib, yB); }        // shared by every template with this B shape
          if (ia >= 0) {
            const double sc = sInorm[g - gb];
#pragma unroll
            for (int j = 0; j < KT; j++) zc[j] = sc * (ysm[((ia - ia0) * KT + j) * FAST_TPB + tid] + yB[j]);
          } else {
#pragma unroll
            for (int j = 0; j < KT; j++) zc[j] = yB[j];
          }
        } else {
          correlate(g - gb, zc);
        }
        const bool half = a.hyp_half[h] != 0;
        const double kg = sK[g - gb];
        double val[KT];
        if (half) {
          bool far_ = false;
#pragma unroll
          for (int j = 0; j < KT; j++) {
            val[j] = kg + phi_tab(zc[j], sphi);
            far_ = far_ || (zc[j] < BFL_PHI_ZFAR);
          }
          if (far_) {   // rare: a > 12 sigma negative amplitude
#pragma unroll
            for (int j = 0; j < KT; j++)
              if (zc[j] < BFL_PHI_ZFAR) val[j] = kg + phi_half_far(zc[j]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < KT; j++) val[j] = fma(0.5 * zc[j], zc[j], kg);
        }
        if (MODE != 1) {
#pragma unroll
          for (int j = 0; j < KT; j++) lse_push(mx[j], sm[j], val[j], st64);
        } else {
          const double hc = (half ? 0.5 * (BFL_LNPI - BFL_LN2) : 0.5 * (BFL_LN2 + BFL_LNPI)) + hlv;
#pragma unroll
          for (int j = 0; j < KT; j++) {
            if (valid[j]) {
              const int64_t ki = kfirst + j - a.k0;
              a.out_full[(size_t)a.gorig[g] * nk + ki] = (val[j] + hc) + (a.bgabs[(size_t)b * nk + ki] + a.sls);
            }
          }
        }
      }
    }
    flush(h);
  }
  if (MODE == 2) {
    double* o = a.lnO + (size_t)b * nk;
#pragma unroll
    for (int j = 0; j < KT; j++)
      if (valid[j]) o[kfirst + j - a.k0] = (n_m[j] - d_m[j]) + log(n_s[j] / d_s[j]);
  }
  }   // warp-tile loop
}

// ======================================================================================
// FAST path, split form for large batches (W = 55, separable-grid plans)
// ======================================================================================
// The correlation wants a big register window (few resident warps); the epilogue (table look-ups,
// exp, log-sum-exp) is latency/LSU bound and wants many resident warps.  One kernel cannot have both,
// so large batches run two:
//   k_corr_shapes : y_s(k) = shape_s . d[k-h .. k+h] for the S shapes of the plan -> Y[s][b*nk + k]
//   k_epilogue    : one sample per thread lane (SPT samples per thread), z_g = inorm_g (yA + yB),
//                   phi table, online log-sum-exp per hypothesis, hypotheses combined -> lnO
// Y makes one round trip through HBM (S * 8 bytes per sample each way), overlapped with the FP64 work.
struct CorrArgs {
  const double* data;
  const double* sk; const double* cle;      // per-curve noise sigma, flux errors (constant per curve here) or nullptr
  double* sqrtu; double* halflogv;          // [B] written for the epilogue (what k_prep_scalars computes)
  unsigned long long* done_counter;         // warps that have left the tile loop: the last one re-arms both counters
  int64_t seglen, N, seg0, k0, k1;
  int B, tiles_per_curve;
  int64_t total_wtiles;
  unsigned long long* tile_counter;
  const double* shapes; int S, WP;
  double* Y;            // [S][B * nk]
};

// W is a compile-time window length (the register window holds W + KT - 1 samples): instantiated for the odd
// lengths the reference's scripts use (21, 33, 55) and 101; other lengths take the single-kernel path.
template <int W, int KT>
__global__ void __launch_bounds__(FAST_TPB, (W + KT - 1 > 64) ? 2 : 3) k_corr_shapes(const CorrArgs a) {
  extern __shared__ double smem[];
  constexpr int H = W / 2, TKW = 32 * KT;
  const int WP = a.WP;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int wstride = (TKW + 2 * H + 2) & ~1;
  double* swin = smem + warp * wstride;
  double* sshape = smem + (FAST_TPB / 32) * wstride;       // [S][WP]
  {
    const double2* src = reinterpret_cast<const double2*>(a.shapes);
    double2* dst = reinterpret_cast<double2*>(sshape);
    for (int i = tid; i < a.S * WP / 2; i += FAST_TPB) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  const int64_t nk = a.k1 - a.k0;
  const int64_t BK = (int64_t)a.B * nk;
  for (;;) {
    long long wt = 0;
    if (lane == 0) wt = (long long)atomicAdd(a.tile_counter, 1ULL);
    wt = __shfl_sync(0xffffffffu, wt, 0);
    if (wt >= a.total_wtiles) {
      // self-resetting scheduler: the last warp of the grid to leave puts both counters back to zero
      if (lane == 0 && atomicAdd(a.done_counter, 1ULL) == (unsigned long long)gridDim.x * (FAST_TPB / 32) - 1ULL) {
        *a.tile_counter = 0ULL;
        *a.done_counter = 0ULL;
      }
      break;
    }
    const int tile = (int)(wt % a.tiles_per_curve);
    const int b = (int)(wt / a.tiles_per_curve);
    const int64_t kbase = a.k0 + (int64_t)tile * TKW;
    {
      // per-curve scalars inline (no separate preparation launch): v = sk^2 + cle^2 as NumPy rounds it (bayes.py:312)
      const double sg = __ldg(a.sk + b), er = a.cle ? __ldg(a.cle + (size_t)b * a.seglen) : 0.0;
      const double var = __dadd_rn(__dmul_rn(sg, sg), __dmul_rn(er, er));
      const double su = sqrt(1.0 / var);
      if (tile == 0 && lane == 0) { a.sqrtu[b] = su; a.halflogv[b] = 0.5 * log(var); }
      const double* d = a.data + (size_t)b * a.seglen;
      const int64_t lo = a.seg0, hi = min(a.N, a.seg0 + a.seglen);
      __syncwarp();
      for (int i = lane; i < TKW + 2 * H; i += 32) {
        const int64_t sidx = kbase - H + i;
        double v = 0.0;
        if (sidx >= lo && sidx < hi) v = __ldg(d + (sidx - a.seg0)) * su;
        swin[i] = v;
      }
      __syncwarp();
    }
    double win[W + KT - 1];
#pragma unroll
    for (int e = 0; e < W + KT - 1; e++) win[e] = swin[lane * KT + e];
    const int64_t kfirst = kbase + (int64_t)lane * KT;
    bool valid[KT];
#pragma unroll
    for (int j = 0; j < KT; j++) {
      const int64_t k = kfirst + j;
      valid[j] = (k < a.k1) && (k >= H) && (k < a.N - H);
    }
    double* yout = a.Y + (size_t)b * nk + (kfirst - a.k0);
    for (int sh = 0; sh < a.S; sh++) {
      double y[KT];
      correlate_reg<W, KT>(sshape + (size_t)sh * WP, win, y);
#pragma unroll
      for (int j = 0; j < KT; j++)
        if (valid[j]) yout[(size_t)sh * BK + j] = y[j];
    }
  }
}

struct EpiArgs {
  const double* Y;        // [S][BK]
  const double* halflogv; // [B]
  const double* sqrtu;    // [B]
  unsigned long long* tile_counter;   // k_epilogue2: next warp tile (self-resetting)
  unsigned long long* done_counter;   // k_epilogue2: warps that have left the tile loop
  const unsigned char* plan_block;    // k_epilogue2: per-plan shared-memory block built by k_epilogue2_prepare
  int srows;              // k_epilogue2: rows of the per-thread slice (A shapes of separable + all shapes of dense hypotheses)
  int64_t nk, BK, N, k0;
  int B, G, nhyp, H;
  int hyp_begin[MAXH + 1];
  int hyp_half[MAXH];
  int hyp_A0[MAXH];
  int hyp_nA[MAXH];
  const double* Kg; const double* lp; const double* lw; const double* inorm;
  const short* idxA; const short* idxB;
  const int* sh_g0; const int* sh_g1;   // per shape: the run of templates that use it as their B shape
  int hyp_B0[MAXH];
  int hyp_nB[MAXH];
  int S, arows;           // shapes; A-shape rows of the per-thread shared-memory slice
  int noisepoly;
  double* lnO;            // MODE 2
  double2* part;          // MODE 0: [nhyp][BK]
};

constexpr int EPI_TPB = 256;
#ifndef BFL_EPI_SPT
#define BFL_EPI_SPT 2      // samples per thread of the epilogue kernel
#endif
#ifndef BFL_EPI_MINB
#define BFL_EPI_MINB 3
#endif
#ifndef BFL_EPI_GU
#define BFL_EPI_GU 1
#endif
constexpr int EPI_GU = BFL_EPI_GU;     // templates evaluated together by the linear-domain loop (2 measured no faster)

// Two evaluation strategies per thread tile:
//  * LINEAR (first): nearly every sample is noise, where every |z_g| is a few units.  Then the grid sum
//    is accumulated directly, S_h = sum_g c_g E(z_g) with c_g = exp(K_g + lp_g + lw_g - Cref_h) fixed per
//    plan and E = exp(phi) from exphi_eval (one 16-byte look-up + ~20 FMAs per template) -- no exp range reduction, no
//    running maximum.  Full-range templates use exp(z^2/2 - ZC^2/2) (one exp, no maximum).
//  * LOG (fallback, only for threads that met a |z| > ZC, i.e. a flare or a deep dip): phi table +
//    online log-sum-exp, any z.  Its phi table is read from global memory (L1): it is rarely needed and
//    shared memory is better spent on occupancy.
// Both produce (reference value, sum) pairs per hypothesis, folded into the odds the same way.
template <int MODE, int SPT>
__global__ void __launch_bounds__(EPI_TPB, BFL_EPI_MINB) k_epilogue(const EpiArgs a) {
  extern __shared__ double smem[];
  const int tid = threadIdx.x;
  const int G = a.G;
  double* sK = smem;                                     // [G] Kg + lp + lw
  double* sC = sK + ((G + 1) & ~1);                      // [G] exp(sK - Cref_h)
  double* sInorm = sC + ((G + 1) & ~1);                  // [G]
  double* sCref = sInorm + ((G + 1) & ~1);               // [MAXH]
  double* sE = sCref + MAXH;                             // E(z) table
  double* ysm = sE + BFL_EXPHI_SIZE;                     // [arows][SPT][EPI_TPB] this thread's A-shape correlations
  int* sG01 = reinterpret_cast<int*>(ysm + (size_t)a.arows * SPT * EPI_TPB);   // [S][2]
  short* sIdxA = reinterpret_cast<short*>(sG01 + 2 * a.S);                     // [G]
  for (int i = tid; i < G; i += EPI_TPB) {
    sK[i] = a.Kg[i] + a.lp[i] + a.lw[i];
    sInorm[i] = a.inorm[i];
    sIdxA[i] = a.idxA[i];
  }
  for (int i = tid; i < a.S; i += EPI_TPB) { sG01[2 * i] = a.sh_g0[i]; sG01[2 * i + 1] = a.sh_g1[i]; }
  for (int i = tid; i < BFL_EXPHI_SIZE; i += EPI_TPB) sE[i] = BFL_EXPHI_TAB[i];
  __syncthreads();
  if (tid < a.nhyp) {
    double m = -1.0e300;
    for (int g = a.hyp_begin[tid]; g < a.hyp_begin[tid + 1]; g++) m = fmax(m, sK[g]);
    sCref[tid] = m;
  }
  __syncthreads();
  for (int i = tid; i < G; i += EPI_TPB) {
    int h = 0;
    while (a.hyp_begin[h + 1] <= i) h++;
    sC[i] = exp(sK[i] - sCref[h]);
  }
  __syncthreads();
  constexpr double ZC = BFL_EXPHI_ZC;

  const int64_t tile_samples = (int64_t)EPI_TPB * SPT;
  const int64_t ntiles = (a.BK + tile_samples - 1) / tile_samples;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int64_t idx[SPT];
    bool valid[SPT];
    double hlv[SPT];
#pragma unroll
    for (int j = 0; j < SPT; j++) {
      idx[j] = tile * tile_samples + (int64_t)j * EPI_TPB + tid;
      valid[j] = idx[j] < a.BK;
      if (!valid[j]) idx[j] = a.BK - 1;
      const int64_t b = idx[j] / a.nk;
      const int64_t k = a.k0 + (idx[j] - b * a.nk);
      valid[j] = valid[j] && (k >= a.H) && (k < a.N - a.H);   // interior samples only
      hlv[j] = a.halflogv[b];
    }
    double lnO[SPT];
    bool big = false;

    auto body = [&](auto LIN) {
      constexpr bool lin = decltype(LIN)::value;
      double mx[SPT], sm[SPT], n_m[SPT], n_s[SPT], d_m[SPT], d_s[SPT];
#pragma unroll
      for (int j = 0; j < SPT; j++) {
        mx[j] = LSE_EMPTY; sm[j] = 0.0;
        n_m[j] = LSE_EMPTY; n_s[j] = 0.0;
        d_m[j] = a.noisepoly ? 0.0 : LSE_EMPTY; d_s[j] = a.noisepoly ? 1.0 : 0.0;
      }
      for (int h = 0; h < a.nhyp; h++) {
        const int glo = a.hyp_begin[h], ghi = a.hyp_begin[h + 1];
        if (glo >= ghi) continue;
        const bool half = a.hyp_half[h] != 0;
        {
          // The hypothesis' A-shape correlations are fetched once (coalesced, all in flight together) into
          // this thread's shared-memory slice; B shapes are walked in order, the next one prefetched
          // while the run of templates that share the current one is evaluated.
          const int nA = a.hyp_nA[h], B0 = a.hyp_B0[h], nB = a.hyp_nB[h];
          {
            const double* YA = a.Y + (size_t)a.hyp_A0[h] * a.BK;
            for (int ia = 0; ia < nA; ia++) {
#pragma unroll
              for (int j = 0; j < SPT; j++) ysm[(ia * SPT + j) * EPI_TPB + tid] = __ldg(YA + (size_t)ia * a.BK + idx[j]);
            }
          }
          double yB[SPT], yBn[SPT];
          const double* YB = a.Y + (size_t)B0 * a.BK;
#pragma unroll
          for (int j = 0; j < SPT; j++) yBn[j] = __ldg(YB + idx[j]);
          for (int ib = 0; ib < nB; ib++) {
            const double* YBn = YB + (size_t)min(ib + 1, nB - 1) * a.BK;
#pragma unroll
            for (int j = 0; j < SPT; j++) { yB[j] = yBn[j]; yBn[j] = __ldg(YBn + idx[j]); }
            const int g0 = sG01[2 * (B0 + ib)], g1 = sG01[2 * (B0 + ib) + 1];
            if constexpr (lin) {
              // EPI_GU templates per iteration: EPI_GU * SPT independent dependency chains per thread
              for (int g = g0; g < g1; g += EPI_GU) {
                double z[EPI_GU][SPT], cg[EPI_GU];
#pragma unroll
                for (int u = 0; u < EPI_GU; u++) {
                  const int gu = min(g + u, g1 - 1);
                  cg[u] = (g + u < g1) ? sC[gu] : 0.0;
                  if (nA > 0) {
                    const double sc = sInorm[gu];
                    const double* ya = ysm + (size_t)sIdxA[gu] * SPT * EPI_TPB + tid;
#pragma unroll
                    for (int j = 0; j < SPT; j++) z[u][j] = sc * (ya[j * EPI_TPB] + yB[j]);
                  } else {
#pragma unroll
                    for (int j = 0; j < SPT; j++) z[u][j] = yB[j];
                  }
                }
#pragma unroll
                for (int u = 0; u < EPI_GU; u++) {
#pragma unroll
                  for (int j = 0; j < SPT; j++) {
                    big = big || ((__double2hiint(z[u][j]) & 0x7fffffff) >= 0x40180000);   // |z| >= 6 (or NaN)
                    const double e = half ? exphi_eval(z[u][j], sE)
                                          : exp_nonpos_poly(fma(0.5 * z[u][j], z[u][j], -0.5 * ZC * ZC));
                    sm[j] = fma(cg[u], e, sm[j]);
                  }
                }
              }
            } else {
            for (int g = g0; g < g1; g++) {
              double z[SPT];
              if (nA > 0) {
                const double sc = sInorm[g];
                const double* ya = ysm + (size_t)sIdxA[g] * SPT * EPI_TPB + tid;
#pragma unroll
                for (int j = 0; j < SPT; j++) z[j] = sc * (ya[j * EPI_TPB] + yB[j]);
              } else {
#pragma unroll
                for (int j = 0; j < SPT; j++) z[j] = yB[j];
              }
              if constexpr (!lin) {
                const double kg = sK[g];
                double val[SPT];
                if (half) {
                  bool far_ = false;
#pragma unroll
                  for (int j = 0; j < SPT; j++) {
                    val[j] = kg + phi_tab_rows(z[j], BFL_PHI_TAB);
                    far_ = far_ || (z[j] < BFL_PHI_ZFAR);
                  }
                  if (far_) {
#pragma unroll
                    for (int j = 0; j < SPT; j++)
                      if (z[j] < BFL_PHI_ZFAR) val[j] = kg + phi_half_far(z[j]);
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < SPT; j++) val[j] = fma(0.5 * z[j], z[j], kg);
                }
#pragma unroll
                for (int j = 0; j < SPT; j++) lse_push_poly(mx[j], sm[j], val[j]);
              }
            }
            }   // log-domain loop
          }
        }
        // hypothesis finished: partial out (MODE 0) or fold into numerator / denominator (MODE 2)
        const double hcb = half ? 0.5 * (BFL_LNPI - BFL_LN2) : 0.5 * (BFL_LN2 + BFL_LNPI);
        if constexpr (lin) {
          const double ref = sCref[h] + (half ? 0.0 : 0.5 * ZC * ZC);
#pragma unroll
          for (int j = 0; j < SPT; j++) mx[j] = ref;
        }
#pragma unroll
        for (int j = 0; j < SPT; j++) {
          const double x = mx[j] + (hcb + hlv[j]);
          if (MODE == 0) {
            if (valid[j]) a.part[(size_t)h * a.BK + idx[j]] = make_double2(x, sm[j]);
          } else if (h == 0) {
            n_m[j] = x; n_s[j] = sm[j];
          } else {
            const double d = x - d_m[j];
            const double e = exp_nonpos_poly(-fabs(d));
            const bool up = d > 0.0;
            d_s[j] = up ? fma(d_s[j], e, sm[j]) : fma(sm[j], e, d_s[j]);
            d_m[j] = up ? x : d_m[j];
          }
          mx[j] = LSE_EMPTY; sm[j] = 0.0;
        }
      }
      if (MODE == 2) {
#pragma unroll
        for (int j = 0; j < SPT; j++) lnO[j] = (n_m[j] - d_m[j]) + log(n_s[j] / d_s[j]);
      }
    };
    body(std::true_type{});
    if (big) body(std::false_type{});
    if (MODE == 2) {
#pragma unroll
      for (int j = 0; j < SPT; j++)
        if (valid[j]) a.lnO[idx[j]] = lnO[j];
    }
  }
}

// --------------------------------------------------------------------------------------
// k_epilogue2: the same stage with the E(z) evaluation in mixed precision (tools/gen_exphi2_table.py).
// ncu of k_epilogue (round 1): 56 warp instructions per (32 samples, template), 28 of them FP64, two
// F2I/I2F on the XU pipe, FP64 pipe 49 % busy.  Only the part of E(z) that needs 53 bits stays in fp64
// (10 instructions): E_k (1 + t + t^2/2) with t = (z - z_k) phi'(z_k).  The remainder of the series
// (< 1e-5 of E) is evaluated in fp32 for the thread's two samples at once with the packed fp32x2
// instructions of sm_100 (FFMA2 / FMUL2 / FADD2) and accumulated in a float per sample; fp32 copies
// of the table row are made by integer truncation (two ALU instructions each), of z - z_k by one F2F.
// A table row is (E_k, phi'(z_k)): still ONE 16-byte lane-varying load per evaluation.
struct __align__(16) EpiTC {   // per-template constants, one 32-byte broadcast record
  double inorm;   // 1/|m_perp|
  double c;       // exp(K_g + lp_g + lw_g - Cref_h)
  float cf;       // (float)c
  int aoff;       // offset of the template's A-shape row in the thread slice (idxA * SPT * EPI_TPB)
  int pad0, pad1;
};


// positive normal float -> double, exact: three integer instructions
__device__ __forceinline__ double f32pos_to_f64(float x) {
  const unsigned b = (unsigned)__float_as_int(x);
  return __hiloint2double((int)((b >> 3) + 0x38000000u), (int)(b << 29));
}
// positive, normal, float-range double -> float by truncation: two integer instructions
__device__ __forceinline__ float f64pos_to_f32_trunc(double x) {
  return __int_as_float(__funnelshift_l(__double2loint(x), __double2hiint(x) - 0x38000000, 3));
}

__device__ __forceinline__ float rcp_approx(float x) {   // one MUFU.RCP, no denormal slow path
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}


__device__ __forceinline__ void exphi2_pair(double z6a, double z6b, unsigned tab, double c0, double c1, double& S,
                                            unsigned& kmax) {
  const double ea = exphi2_one(z6a, tab, kmax), eb = exphi2_one(z6b, tab, kmax);
  S = fma(c0, ea, S);
  S = fma(c1, eb, S);
}

// two templates evaluated together: a pair inside the run of templates that share one B shape (separable
// hypothesis), or two consecutive templates of a dense hypothesis.  An odd one out is paired with a copy of
// itself of weight zero.
struct __align__(16) EpiTP {
  double inorm0, inorm1;
  double c0, c1;
  float cf0, cf1;
  int aoff0, aoff1;       // byte offsets of the templates' A-shape rows in the thread slice
};

constexpr int EPI2_TPB = 256;   // one sample per thread: 64 registers, four CTAs (32 warps) per SM
#ifndef BFL_EPI2_MINB
#define BFL_EPI2_MINB 4
#endif

__device__ __forceinline__ long long f64_order_key(double x) {   // monotone map double -> signed 64-bit
  const long long b = __double_as_longlong(x);
  return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}
__device__ __forceinline__ double f64_from_order_key(long long k) {
  return __longlong_as_double(k ^ ((k >> 63) & 0x7fffffffffffffffLL));
}

// Log odds of one sample per thread from the shape correlations Y (fused detector mode only).
//
// LINEAR pass (every thread): nearly every sample is noise, where all |z| are a few units, so the grid sums are
// accumulated as plain numbers.  The per-template weights carry every per-plan constant, including the
// hypothesis' own normalisation, so all noise hypotheses add into ONE denominator sum and no exp / log-sum-exp
// fold is needed between hypotheses:
//     num = sum_{g in signal}  exp(K_g - M0)                    E(z_g)
//     den = sum_{g in noise h} exp(K_g + hc_h [+ ZC^2/2] - M)   E_h(z_g)      (+ exp(-M) / sqrt(v) for the polynomial-only model)
//     lnO = (M0 + hc_0 - M) + log(num / den)
// with E_h = exp(phi) for half-range amplitudes (exphi2_pair) and exp(z^2/2 - ZC^2/2) for full-range ones.
// LOG pass (only threads that met |z| > ZC, i.e. a flare or a deep dip): phi table + online log-sum-exp per
// hypothesis, any z, as k_epilogue does.
// shared-memory layout of k_epilogue2: [per-plan block, identical in every CTA and every launch][per-thread slice]
struct Epi2Layout {
  int oT, oTP, oTC, oK, oCref, oConst, oG01, oRun, oRow, plan_bytes;   // byte offsets; plan_bytes is a multiple of 16
  __host__ __device__ Epi2Layout(int G, int S) {
    int o = 0;
    oT = o; o += 8 * (BFL_EXPHI2_NI + 1);
    oTP = o; o += (int)sizeof(EpiTP) * (G / 2 + S + 1);
    oTC = o; o += (int)sizeof(EpiTC) * G;
    oK = o; o += 8 * ((G + 1) & ~1);
    oCref = o; o += 8 * MAXH;
    oConst = o; o += 8 * 2;
    oG01 = o; o += 8 * S;
    oRun = o; o += 8 * S;
    oRow = o; o += 4 * S;
    plan_bytes = (o + 15) & ~15;
  }
};

// Builds the per-plan block ONCE per (plan, noisepoly) -- at plan creation, not in every CTA of every launch
// (at 128 light curves per GPU the 592 CTA prologues were a tenth of the kernel): one CTA, result in global memory.
__global__ void __launch_bounds__(EPI2_TPB) k_epilogue2_prepare(const EpiArgs a, unsigned char* __restrict__ blob) {
  extern __shared__ double smem[];
  unsigned char* sb = reinterpret_cast<unsigned char*>(smem);
  const int tid = threadIdx.x;
  const int G = a.G;
  const Epi2Layout L(G, a.S);
  double* sT = reinterpret_cast<double*>(sb + L.oT);                   // [NI + 1] E_k
  EpiTP* sTP = reinterpret_cast<EpiTP*>(sb + L.oTP);                   // [G / 2 + S + 1]
  EpiTC* sTC = reinterpret_cast<EpiTC*>(sb + L.oTC);                   // [G]  (full-range and log-domain loops)
  double* sK = reinterpret_cast<double*>(sb + L.oK);                   // [G] Kg + lp + lw
  double* sCref = reinterpret_cast<double*>(sb + L.oCref);             // [MAXH] per-hypothesis maximum of sK (log pass)
  double* sConst = reinterpret_cast<double*>(sb + L.oConst);           // [2] lnO offset, weight of the polynomial-only model
  int* sG01 = reinterpret_cast<int*>(sb + L.oG01);                     // [S][2] run of templates per B shape
  int2* sRun = reinterpret_cast<int2*>(sb + L.oRun);                   // [S] (byte offset of the run's first pair, pairs)
  int* sRow = reinterpret_cast<int*>(sb + L.oRow);                     // [S] row of shape s in the thread slice, or -1
  long long* sKey = reinterpret_cast<long long*>(sb + L.plan_bytes);   // [MAXH + 2] maxima under construction
  constexpr double ZC = BFL_EXPHI2_ZC;
  for (int i = tid; i < L.plan_bytes / 8; i += EPI2_TPB) smem[i] = 0.0;
  __syncthreads();
  for (int i = tid; i < G; i += EPI2_TPB) sK[i] = a.Kg[i] + a.lp[i] + a.lw[i];
  for (int i = tid; i < a.S; i += EPI2_TPB) { sG01[2 * i] = a.sh_g0[i]; sG01[2 * i + 1] = a.sh_g1[i]; }
  for (int i = tid; i < BFL_EXPHI2_NI; i += EPI2_TPB) sT[i] = BFL_EXPHI2_TAB[i];
  if (tid < MAXH + 2) sKey[tid] = (tid == MAXH && a.noisepoly) ? f64_order_key(0.0) : f64_order_key(-1.0e300);
  __syncthreads();
  for (int i = tid; i < G; i += EPI2_TPB) {
    int h = 0;
    while (a.hyp_begin[h + 1] <= i) h++;
    atomicMax(sKey + h, f64_order_key(sK[i]));
    if (h > 0) {
      const double hc = a.hyp_half[h] ? 0.5 * (BFL_LNPI - BFL_LN2) : 0.5 * (BFL_LN2 + BFL_LNPI) + 0.5 * ZC * ZC;
      atomicMax(sKey + MAXH, f64_order_key(sK[i] + hc));
    }
  }
  if (tid == 64) {   // rows of the thread slice: A shapes of separable hypotheses, every shape of dense ones
    int row = 0;
    for (int sidx = 0; sidx < a.S; sidx++) sRow[sidx] = -1;
    for (int h = 0; h < a.nhyp; h++) {
      if (a.hyp_nA[h] > 0) for (int ia = 0; ia < a.hyp_nA[h]; ia++) sRow[a.hyp_A0[h] + ia] = row++;
      else for (int ib = 0; ib < a.hyp_nB[h]; ib++) sRow[a.hyp_B0[h] + ib] = row++;
    }
  }
  if (tid == 32) {   // pair numbering: runs of a separable hypothesis one after the other, a dense hypothesis as one run
    int cnt = 0;
    for (int h = 0; h < a.nhyp; h++) {
      if (a.hyp_nA[h] > 0) {
        for (int sidx = a.hyp_B0[h]; sidx < a.hyp_B0[h] + a.hyp_nB[h]; sidx++) {
          const int np = (sG01[2 * sidx + 1] - sG01[2 * sidx] + 1) / 2;
          sRun[sidx] = make_int2(48 * cnt, np);
          cnt += np;
        }
      } else {
        const int np = (a.hyp_nB[h] + 1) / 2;
        for (int sidx = a.hyp_B0[h]; sidx < a.hyp_B0[h] + a.hyp_nB[h]; sidx++) sRun[sidx] = make_int2(48 * cnt, np);
        cnt += np;
      }
    }
  }
  __syncthreads();
  if (tid < a.nhyp) sCref[tid] = f64_from_order_key(sKey[tid]);
  if (tid == 0) {
    const double M0 = f64_from_order_key(sKey[0]), M = f64_from_order_key(sKey[MAXH]);
    const double hc0 = a.hyp_half[0] ? 0.5 * (BFL_LNPI - BFL_LN2) : 0.5 * (BFL_LN2 + BFL_LNPI) + 0.5 * ZC * ZC;
    sConst[0] = (M0 + hc0) - M;
    sConst[1] = a.noisepoly ? exp(-M) : 0.0;
  }
  for (int i = tid; i < G; i += EPI2_TPB) {
    int h = 0;
    while (a.hyp_begin[h + 1] <= i) h++;
    const double hc = a.hyp_half[h] ? 0.5 * (BFL_LNPI - BFL_LN2) : 0.5 * (BFL_LN2 + BFL_LNPI) + 0.5 * ZC * ZC;
    const double Mh = (h == 0) ? f64_from_order_key(sKey[0]) : f64_from_order_key(sKey[MAXH]) - hc;
    auto weight = [&](int g) { return exp(sK[g] - Mh); };
    EpiTC tc;
    tc.inorm = a.inorm[i];
    tc.c = weight(i);
    tc.cf = (float)tc.c;
    tc.aoff = (int)a.idxA[i] * EPI2_TPB;
    tc.pad0 = tc.pad1 = 0;
    sTC[i] = tc;
    // the pair record this template opens (even position in its run)
    const int sB = a.idxB[i];
    const int run0 = (a.hyp_nA[h] > 0) ? sG01[2 * sB] : a.hyp_begin[h];
    const int run1 = (a.hyp_nA[h] > 0) ? sG01[2 * sB + 1] : a.hyp_begin[h + 1];
    const int loc = i - run0;
    if ((loc & 1) == 0) {
      EpiTP tp;
      const bool has2 = i + 1 < run1;
      const int i2 = has2 ? i + 1 : i;
      const double c2 = has2 ? weight(i2) : 0.0;
      tp.inorm0 = tc.inorm; tp.inorm1 = a.inorm[i2];
      tp.c0 = tc.c; tp.c1 = c2;
      tp.cf0 = tc.cf; tp.cf1 = (float)c2;
      tp.aoff0 = 8 * tc.aoff; tp.aoff1 = 8 * (int)a.idxA[i2] * EPI2_TPB;
      sTP[sRun[sB].x / 48 + loc / 2] = tp;
    }
  }
  __syncthreads();
  for (int i = tid; i < L.plan_bytes / 16; i += EPI2_TPB)
    reinterpret_cast<uint4*>(blob)[i] = reinterpret_cast<const uint4*>(sb)[i];
}

__global__ void __launch_bounds__(EPI2_TPB, BFL_EPI2_MINB) k_epilogue2(const EpiArgs a) {
  extern __shared__ double smem[];
  unsigned char* sb = reinterpret_cast<unsigned char*>(smem);
  const int tid = threadIdx.x;
  const Epi2Layout L(a.G, a.S);
  for (int i = tid; i < L.plan_bytes / 16; i += EPI2_TPB)
    reinterpret_cast<uint4*>(sb)[i] = __ldg(reinterpret_cast<const uint4*>(a.plan_block) + i);
  double* sT = reinterpret_cast<double*>(sb + L.oT);
  EpiTP* sTP = reinterpret_cast<EpiTP*>(sb + L.oTP);
  EpiTC* sTC = reinterpret_cast<EpiTC*>(sb + L.oTC);
  double* sK = reinterpret_cast<double*>(sb + L.oK);
  double* sConst = reinterpret_cast<double*>(sb + L.oConst);
  int* sG01 = reinterpret_cast<int*>(sb + L.oG01);
  int2* sRun = reinterpret_cast<int2*>(sb + L.oRun);
  int* sRow = reinterpret_cast<int*>(sb + L.oRow);
  double* ysm = reinterpret_cast<double*>(sb + L.plan_bytes);          // [srows][EPI2_TPB] per-thread slice
  constexpr double ZC = BFL_EXPHI2_ZC;
  __syncthreads();
  const unsigned aT = (unsigned)__cvta_generic_to_shared(sT);
  const unsigned aTP = (unsigned)__cvta_generic_to_shared(sTP);
  const unsigned aRun = (unsigned)__cvta_generic_to_shared(sRun);
  const unsigned aY = (unsigned)__cvta_generic_to_shared(ysm + tid);
  const double lnO_off = sConst[0], w_poly = sConst[1];

  // dynamic scheduling per WARP: a warp tile is 32 consecutive samples; warps never wait for each other
  // (a tile with a flare costs three times a noise tile, so a static split leaves SMs idle at the end)
  const int lane = tid & 31;
  const int64_t nwt = (a.BK + 31) / 32;
  for (;;) {
    long long wt = 0;
    if (lane == 0) wt = (long long)atomicAdd(a.tile_counter, 1ULL);
    wt = __shfl_sync(0xffffffffu, wt, 0);
    if (wt >= nwt) {
      if (lane == 0 && atomicAdd(a.done_counter, 1ULL) == (unsigned long long)gridDim.x * (EPI2_TPB / 32) - 1ULL) {
        *a.tile_counter = 0ULL;      // self-resetting scheduler (see k_corr_shapes)
        *a.done_counter = 0ULL;
      }
      break;
    }
    int64_t idx = wt * 32 + lane;
    bool valid = idx < a.BK;
    if (!valid) idx = a.BK - 1;
    const int bcur = (int)(idx / a.nk);
    {
      const int64_t k = a.k0 + (idx - (int64_t)bcur * a.nk);
      valid = valid && (k >= a.H) && (k < a.N - a.H);   // interior samples only
    }
    const double* Yi = a.Y + idx;
    const size_t BK = (size_t)a.BK;

    // ---- one prefetch phase per tile: the A-shape correlations of every separable hypothesis and the
    // correlations of every dense hypothesis go straight from global into this thread's shared-memory slice
    // (cp.async, no registers); only the B shapes of separable hypotheses are streamed through registers,
    // one run ahead.  Slice row of shape s: a.shrow[s] (or -1).
    __syncwarp();
    for (int sidx = 0; sidx < a.S; sidx++) {
      const int row = sRow[sidx];
      if (row >= 0) cp_async_f64(aY + 8u * EPI2_TPB * (unsigned)row, Yi + (size_t)sidx * BK);
    }
    cp_async_commit();

    // ---------------- linear pass ----------------
    double num = 0.0, den = 0.0;
    unsigned kmax = 0;
    bool big = false;
    bool waited = false;
    for (int h = 0; h < a.nhyp; h++) {
      const int nB = a.hyp_nB[h];
      if (nB <= 0) continue;
      const int nA = a.hyp_nA[h], B0 = a.hyp_B0[h];
      const bool half = a.hyp_half[h] != 0;
      double S = 0.0;
      const double* YB = Yi + (size_t)B0 * BK;
      if (half && nA > 0) {
        // B shapes in order, the next one prefetched while the pairs of the current run are evaluated
        double yBn = __ldg(YB);
        if (!waited) { cp_async_wait_all(); waited = true; }
        const unsigned ayh = aY + 8u * EPI2_TPB * (unsigned)sRow[a.hyp_A0[h]];
        unsigned ar = aRun + 8u * (unsigned)B0;
        for (int ib = 0; ib < nB; ib++, ar += 8u) {
          const double yB = yBn;
          if (ib + 1 < nB) YB += BK;
          yBn = __ldg(YB);
          const int2 run = lds_b32x2(ar);
          unsigned ap = aTP + (unsigned)run.x;
          for (int p = 0; p < run.y; p++, ap += 48u) {
            const double2 in = lds_f64x2(ap);              // (inorm0, inorm1)
            const double2 cc = lds_f64x2(ap + 16u);        // (c0, c1)
            const int2 fa = lds_b32x2(ap + 40u);           // (aoff0, aoff1)
            const double y0 = lds_f64(ayh + (unsigned)fa.x), y1 = lds_f64(ayh + (unsigned)fa.y);
            exphi2_pair(fma(in.x, y0 + yB, ZC), fma(in.y, y1 + yB, ZC), aT, cc.x, cc.y, S, kmax);
          }
        }
      } else if (half && nA == 0) {
        // dense hypothesis: consecutive templates in pairs, each with its own (prefetched) shape correlation
        if (!waited) { cp_async_wait_all(); waited = true; }
        const int2 run = lds_b32x2(aRun + 8u * (unsigned)B0);
        unsigned ap = aTP + (unsigned)run.x;
        unsigned ay = aY + 8u * EPI2_TPB * (unsigned)sRow[B0];
        for (int ib = 0; ib < nB; ib += 2, ap += 48u, ay += 16u * EPI2_TPB) {
          const double z0 = lds_f64(ay), z1 = lds_f64(ay + ((ib + 1 < nB) ? 8u * EPI2_TPB : 0u));
          const double2 cc = lds_f64x2(ap + 16u);
          exphi2_pair(z0 + ZC, z1 + ZC, aT, cc.x, cc.y, S, kmax);
        }
      } else {
        // full-range amplitude: exp(z^2/2 - ZC^2/2), no table (also separable full-range hypotheses, unpaired)
        if (!waited) { cp_async_wait_all(); waited = true; }
        const unsigned ayh = aY + 8u * EPI2_TPB * (unsigned)((nA > 0) ? sRow[a.hyp_A0[h]] : 0);
        for (int ib = 0; ib < nB; ib++) {
          const double yB = (nA > 0) ? __ldg(YB + (size_t)ib * BK) : lds_f64(aY + 8u * EPI2_TPB * (unsigned)sRow[B0 + ib]);
          const int g0 = sG01[2 * (B0 + ib)], g1 = sG01[2 * (B0 + ib) + 1];
          for (int g = g0; g < g1; g++) {
            const EpiTC tc = sTC[g];
            const double z = (nA > 0) ? tc.inorm * (lds_f64(ayh + 8u * (unsigned)tc.aoff) + yB) : yB;
            big = big || ((__double2hiint(z) & 0x7fffffff) >= 0x40180000);   // |z| >= 6 (or NaN)
            S = fma(tc.c, exp_nonpos_poly(fma(0.5 * z, z, -0.5 * ZC * ZC)), S);
          }
        }
      }
      if (h == 0) num = S;
      else den += S;
    }
    if (!waited) cp_async_wait_all();
    big = big || (kmax > (unsigned)(BFL_EXPHI2_NI - 1));   // some |z| > ZC: redo in the log domain
    double lnO = lnO_off + log(num / fma(w_poly, a.sqrtu[bcur], den));

    // ---------------- log-domain pass (rare) ----------------
    if (big) {
      double mx = LSE_EMPTY, sm = 0.0, n_m = LSE_EMPTY, n_s = 0.0;
      double d_m = a.noisepoly ? 0.0 : LSE_EMPTY, d_s = a.noisepoly ? 1.0 : 0.0;
      const double hlv = a.halflogv[bcur];
      for (int h = 0; h < a.nhyp; h++) {
        const int nB = a.hyp_nB[h];
        if (nB <= 0) continue;
        const int nA = a.hyp_nA[h], B0 = a.hyp_B0[h];
        const bool half = a.hyp_half[h] != 0;
        const unsigned ayh = aY + 8u * EPI2_TPB * (unsigned)((nA > 0) ? sRow[a.hyp_A0[h]] : 0);
        for (int ib = 0; ib < nB; ib++) {
          const double yB = __ldg(Yi + (size_t)(B0 + ib) * BK);
          const int g0 = sG01[2 * (B0 + ib)], g1 = sG01[2 * (B0 + ib) + 1];
          for (int g = g0; g < g1; g++) {
            const EpiTC tc = sTC[g];
            const double z = (nA > 0) ? tc.inorm * (lds_f64(ayh + 8u * (unsigned)tc.aoff) + yB) : yB;
            const double kg = sK[g];
            double val;
            if (half) val = kg + ((z < BFL_PHI_ZFAR) ? phi_half_far(z) : phi_tab_rows(z, BFL_PHI_TAB));
            else val = fma(0.5 * z, z, kg);
            lse_push_poly(mx, sm, val);
          }
        }
        const double x = mx + ((half ? 0.5 * (BFL_LNPI - BFL_LN2) : 0.5 * (BFL_LN2 + BFL_LNPI)) + hlv);
        if (h == 0) { n_m = x; n_s = sm; }
        else {
          const double d = x - d_m;
          const double e = exp_nonpos_poly(-fabs(d));
          const bool up = d > 0.0;
          d_s = up ? fma(d_s, e, sm) : fma(sm, e, d_s);
          d_m = up ? x : d_m;
        }
        mx = LSE_EMPTY; sm = 0.0;
      }
      lnO = (n_m - d_m) + log(n_s / d_s);
    }
    if (valid) a.lnO[idx] = lnO;
  }
}

static int epi2_srows(const PlanView& pv) {
  int r = 0;
  for (int h = 0; h < pv.nhyp; h++) r += (pv.hyp_nA[h] > 0) ? pv.hyp_nA[h] : pv.hyp_nB[h];
  return r;
}
static size_t epi2_smem_bytes(int G, int S, int srows) {
  return (size_t)Epi2Layout(G, S).plan_bytes + sizeof(double) * (size_t)srows * EPI2_TPB;
}
bool epi2_fits(const PlanView& pv) {   // one sample per thread: the slice must fit four times per SM
  return pv.sep_ok && pv.wsep_ok && split_window_supported(pv.W) && pv.G > 0 && pv.nhyp >= 1 && pv.hyp_begin[1] > 0 &&
         epi2_smem_bytes(pv.G, pv.S, epi2_srows(pv)) <= 56 * 1024;
}
size_t epi2_plan_block_bytes(const PlanView& pv) { return (size_t)Epi2Layout(pv.G, pv.S).plan_bytes; }

static size_t corr_smem_bytes(int S, int KT, int W = 55) {
  const int wstride = (32 * KT + 2 * (W / 2) + 2) & ~1;
  return sizeof(double) * (size_t)((FAST_TPB / 32) * wstride + (size_t)S * (W + 1));
}
// window lengths the split form is instantiated for, and the samples per thread its register window allows
bool split_window_supported(int W) { return W == 21 || W == 33 || W == 55 || W == 101; }
static int split_max_kt(int W) { return W == 101 ? 2 : 4; }
static size_t epi_smem_bytes(int G, int S, int arows, int SPT) {
  return sizeof(double) * ((size_t)(3 * ((G + 1) & ~1) + MAXH + BFL_EXPHI_SIZE) + (size_t)arows * SPT * EPI_TPB) +
         sizeof(int) * (size_t)(2 * S) + sizeof(short) * (size_t)(G + 8);
}

static int epi_arows(const PlanView& pv) {
  int m = 0;
  for (int h = 0; h < pv.nhyp; h++) m = std::max(m, pv.hyp_nA[h]);
  return m;
}

bool fast_split_fits(const PlanView& pv) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("BFL_NO_SPLIT"); off = (e && atoi(e)) ? 1 : 0; }
  return !off && pv.sep_ok && split_window_supported(pv.W) && corr_smem_bytes(pv.S, 4, pv.W) <= 70 * 1024 && pv.wsep_ok &&
         epi_smem_bytes(pv.G, pv.S, epi_arows(pv), BFL_EPI_SPT) <= 50 * 1024 * BFL_EPI_SPT;   // wsep_ok: templates grouped by B shape
}

static void fill_epi_plan_args(EpiArgs& a, const PlanView& pv, int noisepoly) {
  a.Y = nullptr; a.halflogv = nullptr; a.sqrtu = nullptr; a.tile_counter = nullptr; a.done_counter = nullptr; a.plan_block = nullptr;
  a.nk = a.BK = a.N = a.k0 = 0; a.B = 0; a.lnO = nullptr; a.part = nullptr;
  a.G = pv.G; a.nhyp = pv.nhyp; a.H = pv.W / 2;
  for (int h = 0; h <= MAXH; h++) a.hyp_begin[h] = (h <= pv.nhyp) ? pv.hyp_begin[h] : pv.hyp_begin[pv.nhyp];
  for (int h = 0; h < MAXH; h++) {
    const bool on = h < pv.nhyp;
    a.hyp_half[h] = on ? pv.hyp_half[h] : 0; a.hyp_A0[h] = on ? pv.hyp_A0[h] : 0; a.hyp_nA[h] = on ? pv.hyp_nA[h] : 0;
    a.hyp_B0[h] = on ? pv.hyp_B0[h] : 0; a.hyp_nB[h] = on ? pv.hyp_nB[h] : 0;
  }
  a.Kg = pv.Kg; a.lp = pv.lp; a.lw = pv.lw; a.inorm = pv.inorm; a.idxA = pv.idxA; a.idxB = pv.idxB;
  a.sh_g0 = pv.sh_g0; a.sh_g1 = pv.sh_g1; a.S = pv.S; a.arows = epi_arows(pv); a.srows = epi2_srows(pv);
  a.noisepoly = noisepoly;
}

// per-plan shared-memory block of k_epilogue2 for one value of noisepoly -> d_block (epi2_plan_block_bytes(pv) bytes)
cudaError_t launch_epilogue2_prepare(const PlanView& pv, int noisepoly, void* d_block, cudaStream_t s, LaunchStats* st) {
  EpiArgs a;
  fill_epi_plan_args(a, pv, noisepoly);
  const size_t smem = epi2_plan_block_bytes(pv) + sizeof(long long) * (MAXH + 2);
  cudaError_t e = cudaFuncSetAttribute(k_epilogue2_prepare, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_epilogue2_prepare<<<1, EPI2_TPB, smem, s>>>(a, (unsigned char*)d_block);
  st->launches++;
  return cudaGetLastError();
}

// mode: 0 = partial sums per hypothesis (sc.part, NC = sub = 1 layout), 2 = log odds
cudaError_t launch_window_split(const PlanView& pv, const SeriesView& sv, Scratch& sc, int mode, int noisepoly,
                                double* d_lnO, cudaStream_t s, LaunchStats* st) {
  const int sms = sm_count_current();
  const int64_t nk = sv.k1 - sv.k0;
  const int KT = sc.KT;
  CorrArgs c;
  c.data = sv.clc; c.sk = sv.sk; c.cle = sv.cle; c.sqrtu = sc.sqrtu; c.halflogv = sc.halflogv; c.seglen = sv.seglen; c.N = sv.N; c.seg0 = sv.seg0; c.k0 = sv.k0; c.k1 = sv.k1;
  c.B = sv.B; c.tiles_per_curve = (int)((nk + 32 * KT - 1) / (32 * KT));
  c.total_wtiles = (int64_t)c.tiles_per_curve * sv.B; c.tile_counter = sc.split_counters; c.done_counter = sc.split_counters + 2;
  c.shapes = pv.shapes; c.S = pv.S; c.WP = pv.WP; c.Y = sc.Y;
  const size_t smem1 = corr_smem_bytes(pv.S, KT, pv.W);
  cudaError_t e;
#define BFL_CORR(W_, KT_)                                                                                   \
  do {                                                                                                      \
    e = cudaFuncSetAttribute(k_corr_shapes<W_, KT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1); \
    if (e != cudaSuccess) return e;                                                                         \
    int occ = 1;                                                                                            \
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_corr_shapes<W_, KT_>, FAST_TPB, smem1);       \
    if (e != cudaSuccess) return e;                                                                         \
    int64_t gx = (int64_t)sms * (occ < 1 ? 1 : occ);                                                        \
    const int64_t need = (c.total_wtiles + FAST_TPB / 32 - 1) / (FAST_TPB / 32);                            \
    if (gx > need) gx = need;                                                                               \
    k_corr_shapes<W_, KT_><<<(unsigned)(gx < 1 ? 1 : gx), FAST_TPB, smem1, s>>>(c);                         \
  } while (0)
#define BFL_CORR_W(W_) do { if (KT == 4) BFL_CORR(W_, 4); else if (KT == 2) BFL_CORR(W_, 2); else BFL_CORR(W_, 1); } while (0)
  if (pv.W == 21) BFL_CORR_W(21);
  else if (pv.W == 33) BFL_CORR_W(33);
  else if (pv.W == 101) { if (KT >= 2) BFL_CORR(101, 2); else BFL_CORR(101, 1); }
  else BFL_CORR_W(55);
#undef BFL_CORR_W
#undef BFL_CORR
  st->launches++;
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;

  EpiArgs a;
  fill_epi_plan_args(a, pv, noisepoly);
  a.Y = sc.Y; a.halflogv = sc.halflogv; a.sqrtu = sc.sqrtu; a.tile_counter = sc.split_counters + 1; a.done_counter = sc.split_counters + 3; a.nk = nk; a.BK = (int64_t)sv.B * nk; a.N = sv.N; a.k0 = sv.k0;
  a.B = sv.B; a.lnO = d_lnO; a.part = sc.part;
  constexpr int SPT = BFL_EPI_SPT;
  const size_t smem2 = epi_smem_bytes(pv.G, pv.S, a.arows, SPT);
  const int64_t ntiles = (a.BK + EPI_TPB * SPT - 1) / (EPI_TPB * SPT);
#define BFL_EPI(MODE_)                                                                                      \
  do {                                                                                                      \
    e = cudaFuncSetAttribute(k_epilogue<MODE_, SPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2); \
    if (e != cudaSuccess) return e;                                                                         \
    int occ = 1;                                                                                            \
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_epilogue<MODE_, SPT>, EPI_TPB, smem2);        \
    if (e != cudaSuccess) return e;                                                                         \
    int64_t gx = (int64_t)sms * (occ < 1 ? 1 : occ);                                                        \
    if (gx > ntiles) gx = ntiles;                                                                           \
    k_epilogue<MODE_, SPT><<<(unsigned)(gx < 1 ? 1 : gx), EPI_TPB, smem2, s>>>(a);                          \
  } while (0)
  static int epi_version = -1;   // BFL_EPI=1 selects the round-1 all-fp64 evaluation (kept for A/B measurements)
  if (epi_version < 0) { const char* ev = getenv("BFL_EPI"); epi_version = (ev && atoi(ev) == 1) ? 1 : 2; }
  a.plan_block = (const unsigned char*)pv.epi2_block[noisepoly ? 1 : 0];
  if (epi_version == 2 && mode == 2 && a.plan_block) {
    const size_t smem3 = epi2_smem_bytes(pv.G, pv.S, a.srows);
    e = cudaFuncSetAttribute(k_epilogue2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3);
    if (e != cudaSuccess) return e;
    int occ = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_epilogue2, EPI2_TPB, smem3);
    if (e != cudaSuccess) return e;
    int64_t gx = (int64_t)sms * (occ < 1 ? 1 : occ);
    const int64_t ntiles2 = (a.BK + EPI2_TPB - 1) / EPI2_TPB;   // CTA-sized groups of warp tiles
    if (gx > ntiles2) gx = ntiles2;
    k_epilogue2<<<(unsigned)(gx < 1 ? 1 : gx), EPI2_TPB, smem3, s>>>(a);
    st->launches++;
    return cudaGetLastError();
  }
  if (mode == 2) BFL_EPI(2); else BFL_EPI(0);
#undef BFL_EPI
  st->launches++;
  return cudaGetLastError();
}

size_t fast_smem_bytes(int W, int gchunk, int KT) {
  const int WP = W + 1;
  const int wstride = (32 * KT + 2 * (W / 2) + 2) & ~1;
  return sizeof(double) * (size_t)((FAST_TPB / 32) * wstride + (size_t)(gchunk + 1) * WP + ((gchunk + 1) & ~1) +
                                   BFL_PHI_CM_SIZE + 64);
}

// separable-grid form: S shape rows instead of G template rows, per-template scalars, per-thread A slice
size_t fast_sep_smem_bytes(int W, int G, int S, int KT, int arows) {
  const int WP = W + 1;
  const int wstride = (32 * KT + 2 * (W / 2) + 2) & ~1;
  return sizeof(double) * (size_t)((FAST_TPB / 32) * wstride + (size_t)(S + 1) * WP + 2 * ((G + 1) & ~1) +
                                   BFL_PHI_CM_SIZE + 64 + (size_t)arows * KT * FAST_TPB) +
         sizeof(short) * (size_t)((2 * G + 7) & ~7);
}

// A-shape rows the per-thread shared-memory slice must hold: the largest hypothesis, capped by SEP_MAXA
static int sep_arows(const PlanView& pv) {
  int m = 1;
  for (int h = 0; h < pv.nhyp; h++) m = std::max(m, std::min(pv.hyp_nA[h], SEP_MAXA));
  return m;
}

// choose samples-per-thread and the template chunk so that the grid covers the chip
int pick_gchunk(int G, int W, int64_t total_samples, int sm_count, int* KT_out) {
  const int64_t want = 2LL * sm_count;   // at least two CTAs per SM when the problem allows
  int KT = 4;
  if (const char* e = getenv("BFL_KT")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8) KT = v; }
  if (!split_window_supported(W)) KT = 1;
  else {
    KT = std::min(KT, split_max_kt(W));
    while (KT > 1 && (total_samples + FAST_TPB * KT - 1) / (FAST_TPB * KT) < want) KT >>= 1;
  }
  const int64_t tiles = (total_samples + FAST_TPB * KT - 1) / (FAST_TPB * KT);   // CTA-sized groups of warp tiles
  const int WP = W + 1;
  int gmax = (int)((96 * 1024) / (sizeof(double) * (WP + 1)));   // <= 96 KB of templates per CTA
  if (gmax < 1) gmax = 1;
  int NC = (G + gmax - 1) / gmax;
  if (NC < 1) NC = 1;
  if (tiles * NC < want) {
    int64_t nc = (want + tiles - 1) / tiles;
    if (nc > G) nc = G;
    if (nc > NC) NC = (int)nc;
  }
  if (NC < 1) NC = 1;
  int gchunk = (G + NC - 1) / NC;
  if (gchunk < 1) gchunk = 1;
  *KT_out = KT;
  return gchunk;
}

static void fill_fast_args(FastArgs& a, const PlanView& pv, const SeriesView& sv, const Scratch& sc, int KT) {
  a.data = sv.clc;
  a.sqrtu = sc.sqrtu;
  a.halflogv = sc.halflogv;
  a.seglen = sv.seglen; a.N = sv.N; a.seg0 = sv.seg0; a.k0 = sv.k0; a.k1 = sv.k1;
  a.B = sv.B;
  const int64_t nk = sv.k1 - sv.k0;
  a.tiles_per_curve = (int)((nk + 32 * KT - 1) / (32 * KT));
  a.total_wtiles = (int64_t)a.tiles_per_curve * sv.B;
  a.tile_counters = sc.tile_counters;
  a.hat = pv.hat; a.Kg = pv.Kg; a.lp = pv.lp; a.lw = pv.lw; a.gorig = pv.gorig;
  a.G = pv.G; a.W = pv.W; a.WP = pv.WP; a.gchunk = sc.gchunk; a.NC = sc.NC; a.nhyp = pv.nhyp;
  for (int h = 0; h <= pv.nhyp; h++) a.hyp_begin[h] = pv.hyp_begin[h];
  for (int h = 0; h < pv.nhyp; h++) a.hyp_half[h] = pv.hyp_half[h];
  a.part = sc.part;
  a.out_full = nullptr; a.bgabs = sc.bgabs; a.sls = 0.0; a.ihyp = 0; a.lnO = nullptr; a.noisepoly = 0;
  a.shapes = pv.shapes; a.inorm = pv.inorm; a.idxA = pv.idxA; a.idxB = pv.idxB; a.S = pv.S; a.arows = sep_arows(pv);
  for (int h = 0; h < pv.nhyp; h++) { a.hyp_A0[h] = pv.hyp_A0[h]; a.hyp_nA[h] = pv.hyp_nA[h]; }
}

template <int MODE>
static cudaError_t launch_fast_impl(FastArgs& a, int KT, bool sep, cudaStream_t s, LaunchStats* st) {
  const size_t smem = sep ? fast_sep_smem_bytes(a.W, a.G, a.S, KT, a.arows) : fast_smem_bytes(a.W, a.gchunk, KT);
  const int sms = sm_count_current();
  cudaError_t e = cudaSuccess;
  // persistent grid: as many CTAs as the device holds at once (or fewer if there is less work)
#define BFL_LAUNCH(WT_, KT_, SEP_)                                                                        \
  do {                                                                                                    \
    auto kern = k_window_fast<WT_, KT_, MODE, SEP_>;                                                      \
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);               \
    if (e != cudaSuccess) return e;                                                                       \
    int occ = 1;                                                                                          \
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, FAST_TPB, smem);                        \
    if (e != cudaSuccess) return e;                                                                       \
    if (occ < 1) occ = 1;                                                                                 \
    int64_t gx = ((int64_t)sms * occ + a.NC - 1) / a.NC;                                                  \
    const int64_t need = (a.total_wtiles + FAST_TPB / 32 - 1) / (FAST_TPB / 32);                          \
    if (gx > need) gx = need;                                                                             \
    if (gx < 1) gx = 1;                                                                                   \
    dim3 grid((unsigned)gx, (unsigned)a.NC);                                                              \
    kern<<<grid, FAST_TPB, smem, s>>>(a);                                                                 \
  } while (0)
  if (sep) {
    if (KT == 4) BFL_LAUNCH(55, 4, true);
    else if (KT == 2) BFL_LAUNCH(55, 2, true);
    else BFL_LAUNCH(55, 1, true);
  } else if (a.W == 55 && KT == 8) BFL_LAUNCH(55, 8, false);
  else if (a.W == 55 && KT == 4) BFL_LAUNCH(55, 4, false);
  else if (a.W == 55 && KT == 2) BFL_LAUNCH(55, 2, false);
  else if (a.W == 55 && KT == 1) BFL_LAUNCH(55, 1, false);
  else BFL_LAUNCH(0, 1, false);
#undef BFL_LAUNCH
  st->launches++;
  return cudaGetLastError();
}

// the separable form needs the whole plan in one CTA (NC == 1) and W == 55
bool fast_sep_fits(const PlanView& pv, int KT) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("BFL_NO_SEP"); off = (e && atoi(e)) ? 1 : 0; }
  return !off && pv.sep_ok && pv.W == 55 && KT <= 4 && fast_sep_smem_bytes(pv.W, pv.G, pv.S, KT, sep_arows(pv)) <= 220 * 1024;
}
static bool use_sep(const PlanView& pv, const Scratch& sc, int KT) { return sc.NC == 1 && fast_sep_fits(pv, KT); }

static int kt_for(const PlanView& pv, const SeriesView& sv, const Scratch& sc) {
  (void)sv;
  return pv.W == 55 ? sc.KT : 1;     // the single-kernel path has a register window for W = 55 only
}

cudaError_t launch_window_fast_lse(const PlanView& pv, const SeriesView& sv, Scratch& sc, cudaStream_t s,
                                   LaunchStats* st) {
  if (pv.G == 0) return cudaSuccess;
  FastArgs a;
  const int KT = kt_for(pv, sv, sc);
  fill_fast_args(a, pv, sv, sc, KT);
  return launch_fast_impl<0>(a, KT, use_sep(pv, sc, KT), s, st);
}

// fused detector: log odds straight from the window kernel (no partials, no combine launch).
// Valid when every template sits in one chunk, there is at least one noise model, and the sample
// range holds no truncated windows.
bool fast_can_fuse(const PlanView& pv, const Scratch& sc, int noisepoly) {
  return sc.NC == 1 && sc.sub == 1 && (pv.nhyp - 1 + (noisepoly ? 1 : 0)) > 0 && pv.G > 0 && pv.hyp_begin[1] > 0;
}

cudaError_t launch_window_fast_fused(const PlanView& pv, const SeriesView& sv, Scratch& sc, int noisepoly,
                                     double* d_lnO, cudaStream_t s, LaunchStats* st) {
  FastArgs a;
  const int KT = kt_for(pv, sv, sc);
  fill_fast_args(a, pv, sv, sc, KT);
  a.lnO = d_lnO; a.noisepoly = noisepoly;
  return launch_fast_impl<2>(a, KT, use_sep(pv, sc, KT), s, st);
}

cudaError_t launch_window_fast_full(const PlanView& pv, int ihyp, const SeriesView& sv, Scratch& sc,
                                    double sumlogsigma, double* d_out, cudaStream_t s, LaunchStats* st) {
  if (pv.hyp_begin[ihyp + 1] == pv.hyp_begin[ihyp]) return cudaSuccess;
  FastArgs a;
  const int KT = kt_for(pv, sv, sc);
  fill_fast_args(a, pv, sv, sc, KT);
  a.out_full = d_out; a.sls = sumlogsigma; a.ihyp = ihyp;
  return launch_fast_impl<1>(a, KT, use_sep(pv, sc, KT), s, st);
}

// --------------------------------------------------------------------------------------
// polynomial-only factor, interior constant-variance samples:
//   lnB_poly(k) = -0.5 logdet(G) + P * 0.5 log v + 0.5 P log(2 pi) + 0.5 * sum_j (q_j . d sqrt(u))^2
// (bayes.py:439-563 -> general.pyx:315-365 -> log_marg_amp_full_C with lastHalfRange = 0)
__global__ void __launch_bounds__(128) k_bgproj(const double* __restrict__ data, const double* __restrict__ sqrtu,
                                                const double* __restrict__ halflogv, int64_t seglen, int64_t N,
                                                int64_t seg0, int64_t k0, int64_t k1, int tiles_per_curve,
                                                const double* __restrict__ qb, int W, int WP, int P,
                                                double logdetG, double* __restrict__ bgabs) {
  extern __shared__ double smem[];
  const int H = W / 2, tid = threadIdx.x;
  const int tile = blockIdx.x % tiles_per_curve, b = blockIdx.x / tiles_per_curve;
  const int64_t kbase = k0 + (int64_t)tile * 128;
  const int ndata = 128 + 2 * H;
  double* sdata = smem;
  double* sq = smem + ndata;
  const double su = sqrtu[b];
  const double* d = data + (size_t)b * seglen;
  const int64_t lo = seg0, hi = min(N, seg0 + seglen);
  for (int i = tid; i < ndata; i += 128) {
    const int64_t s = kbase - H + i;
    sdata[i] = (s >= lo && s < hi) ? d[s - seg0] * su : 0.0;
  }
  for (int i = tid; i < P * WP; i += 128) sq[i] = qb[i];
  __syncthreads();
  const int64_t k = kbase + tid;
  if (k >= k1 || k < H || k >= N - H) return;
  double S = 0.0;
  for (int j = 0; j < P; j++) {
    const double* q = sq + j * WP;
    double e0 = 0.0, e1 = 0.0;
    int w = 0;
    for (; w + 1 < W; w += 2) { e0 = fma(q[w], sdata[tid + w], e0); e1 = fma(q[w + 1], sdata[tid + w + 1], e1); }
    if (w < W) e0 = fma(q[w], sdata[tid + w], e0);
    const double e = e0 + e1;
    S = fma(e, e, S);
  }
  const double c = -0.5 * logdetG + P * halflogv[b] + 0.5 * P * (BFL_LN2 + BFL_LNPI);
  bgabs[(size_t)b * (k1 - k0) + (k - k0)] = c + 0.5 * S;
}

cudaError_t launch_bgproj(const PlanView& pv, const SeriesView& sv, Scratch& sc, cudaStream_t s, LaunchStats* st) {
  const int64_t nk = sv.k1 - sv.k0;
  const int tiles = (int)((nk + 127) / 128);
  const size_t smem = sizeof(double) * (size_t)(128 + 2 * (pv.W / 2) + pv.P * pv.WP);
  k_bgproj<<<(unsigned)(tiles * sv.B), 128, smem, s>>>(sv.clc, sc.sqrtu, sc.halflogv, sv.seglen, sv.N, sv.seg0,
                                                      sv.k0, sv.k1, tiles, pv.qb, pv.W, pv.WP, pv.P, pv.logdetG,
                                                      sc.bgabs);
  st->launches++;
  return cudaGetLastError();
}

// ======================================================================================
// GENERAL path: reference-order arithmetic, no FMA contraction
// ======================================================================================
__device__ __forceinline__ double mul_(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double div_(double a, double b) { return __ddiv_rn(a, b); }

// np.sum order over n terms (NumPy pairwise summation, n <= 128 blocks)
template <class F>
__device__ __forceinline__ double pairwise_np(int n, F term) {
  if (n < 8) {
    double r = 0.0;
    for (int i = 0; i < n; i++) r = add_(r, term(i));
    return r;
  }
  double r0 = term(0), r1 = term(1), r2 = term(2), r3 = term(3), r4 = term(4), r5 = term(5), r6 = term(6), r7 = term(7);
  int i = 8;
  const int nb = n - (n % 8);
  for (; i < nb; i += 8) {
    r0 = add_(r0, term(i)); r1 = add_(r1, term(i + 1)); r2 = add_(r2, term(i + 2)); r3 = add_(r3, term(i + 3));
    r4 = add_(r4, term(i + 4)); r5 = add_(r5, term(i + 5)); r6 = add_(r6, term(i + 6)); r7 = add_(r7, term(i + 7));
  }
  double res = add_(add_(add_(r0, r1), add_(r2, r3)), add_(add_(r4, r5), add_(r6, r7)));
  for (; i < n; i++) res = add_(res, term(i));
  return res;
}

// state of the reference's elimination (log_marg_amp_full.c:29-49) after the P background
// amplitudes have been completed-the-square, shared by every template at this sample
template <int P>
struct BgElim {
  double sq[P];        // pivots
  double c[P][P];      // final rows: c[i][i] linear coefficient, c[i][j>i] cross coefficients
  double inv2[P], inv4[P];
  double Zm1;          // Z after steps 0..P-2   (polynomial-only hypothesis stops there)
  double Z;            // Z after steps 0..P-1
  double logsum;       // -0.5 * sum_{i<P} log(sq[i]) accumulated in the reference's order
  bool finite;
};

template <int P>
__device__ __forceinline__ void bg_eliminate(BgElim<P>& e, const double (&M)[P][P], const double (&D)[P]) {
  e.finite = true;
#pragma unroll
  for (int i = 0; i < P; i++) {
    e.sq[i] = M[i][i];
    e.c[i][i] = mul_(-2.0, D[i]);
    if (!isfinite(e.sq[i]) || !isfinite(e.c[i][i])) e.finite = false;
#pragma unroll
    for (int j = i + 1; j < P; j++) {
      e.c[i][j] = mul_(2.0, M[i][j]);
      if (!isfinite(e.c[i][j])) e.finite = false;
    }
  }
  double Z = 0.0;
  e.Zm1 = 0.0;
#pragma unroll
  for (int i = 0; i < P; i++) {
    const double inv = div_(1.0, e.sq[i]);
    e.inv2[i] = mul_(0.5, inv);
    e.inv4[i] = mul_(0.25, inv);
    if (i == P - 1) e.Zm1 = Z;
    Z = sub_(Z, mul_(mul_(e.c[i][i], e.c[i][i]), e.inv4[i]));
#pragma unroll
    for (int k = i + 1; k < P; k++) e.sq[k] = sub_(e.sq[k], mul_(mul_(e.c[i][k], e.c[i][k]), e.inv4[i]));
#pragma unroll
    for (int j = i + 1; j < P; j++) {
      e.c[j][j] = sub_(e.c[j][j], mul_(mul_(e.c[i][j], e.c[i][i]), e.inv2[i]));
#pragma unroll
      for (int k = j + 1; k < P; k++) e.c[j][k] = sub_(e.c[j][k], mul_(mul_(e.c[i][j], e.c[i][k]), e.inv2[i]));
    }
  }
  e.Z = Z;
  double ls = 0.0;
#pragma unroll
  for (int i = 0; i < P; i++) ls = sub_(ls, mul_(0.5, log(e.sq[i])));
  e.logsum = ls;
}

// polynomial-only log Bayes factor: log_marg_amp_full_C(P, ..., sigma=1, lastHalfRange=0)
template <int P>
__device__ __forceinline__ double bg_only_logL(const BgElim<P>& e) {
  if (!e.finite) return neg_inf();
  const double X = e.sq[P - 1], Y = e.c[P - 1][P - 1];
  double logL = e.logsum;
  const double q = div_(mul_(mul_(0.25, Y), Y), X);
  logL = sub_(logL, mul_(0.5, sub_(e.Zm1, q)));
  logL = add_(logL, mul_(mul_(0.5, (double)P), add_(BFL_LN2, BFL_LNPI)));
  return logL;
}

// signal hypothesis: finish the elimination for the template column and evaluate
// log_marg_amp_full_C(P+1, ..., sigma=1, lastHalfRange)
template <int P>
__device__ __forceinline__ double signal_logL(const BgElim<P>& e, const double (&mdbg)[P], double mm, double dm,
                                              bool half) {
  double cs[P];
  bool fin = e.finite && isfinite(mm);
  double sqs = mm;
  double cd = mul_(-2.0, dm);
  fin = fin && isfinite(cd);
#pragma unroll
  for (int i = 0; i < P; i++) {
    cs[i] = mul_(2.0, mdbg[i]);
    fin = fin && isfinite(cs[i]);
  }
  if (!fin) return neg_inf();
#pragma unroll
  for (int i = 0; i < P; i++) {
    sqs = sub_(sqs, mul_(mul_(cs[i], cs[i]), e.inv4[i]));
#pragma unroll
    for (int j = i + 1; j < P; j++) cs[j] = sub_(cs[j], mul_(mul_(e.c[i][j], cs[i]), e.inv2[i]));
    cd = sub_(cd, mul_(mul_(cs[i], e.c[i][i]), e.inv2[i]));
  }
  const double X = sqs, Y = cd;
  double logL = sub_(e.logsum, mul_(0.5, log(X)));
  const double q = div_(mul_(mul_(0.25, Y), Y), X);
  logL = sub_(logL, mul_(0.5, sub_(e.Z, q)));
  const double log2pi = add_(BFL_LN2, BFL_LNPI), logpi_2 = sub_(BFL_LNPI, BFL_LN2);
  if (half) {
    const double arg = div_(mul_(0.5, Y), sqrt(mul_(2.0, X)));
    logL = add_(logL, add_(add_(mul_(mul_(0.5, (double)P), log2pi), mul_(0.5, logpi_2)), log_erfc_dev(arg)));
  } else {
    logL = add_(logL, mul_(mul_(0.5, (double)(P + 1)), log2pi));
  }
  return logL;
}

// --------------------------------------------------------------------------------------
// WHOLE-SERIES background (bglen = None, nsinusoids = 0; stats/bayes.py:247-437, the `else` branches at
// :292-294, 330-331, 394-395, 408-409): ONE polynomial background over all N samples -- its Gram matrix and data
// projections are constants of the series (the caller computes them exactly as the reference does, np.sum) --
// and the full-length template slid along the series: for sample k the template's centre sample m[N/2] sits on k,
// truncated to the series.  Per (template, sample): P + 2 dot products over the overlap, then the reference's
// elimination (log_marg_amp_full.c:29-66) in its own order.  One thread per sample, blockIdx.y = template.
struct WholeArgs {
  const double* tm;      // [G][NP] templates on the series' time stamps
  const double* abg;     // [P][N]  bgmodels[j] / noisevar            (bayes.py:394)
  const double* dt;      // [N]     clc / noisevar                     (bayes.py:403)
  const double* v;       // [N]     noisevar
  const double* bgcross; // [P][P]  upper triangle used               (bayes.py:330-331)
  const double* dbgr;    // [P]                                        (bayes.py:408-409)
  const double* lp;      // [G] log prior + amplitude prior; -inf rows are skipped
  int N, NP, G, half;
  double sls;            // sumlogsigma (general.pyx:216-221)
  double* out;           // [G][N]
  double* out_bg;        // [1] or nullptr: the polynomial-only factor (bayes.py:439-563; constant along the series)
};

template <int P>
__global__ void __launch_bounds__(128) k_whole_series(const WholeArgs a) {
  const int g = blockIdx.y;
  const int k = blockIdx.x * 128 + threadIdx.x;
  const int N = a.N;
  if (k >= N) return;
  double M[P][P], D[P];
#pragma unroll
  for (int i = 0; i < P; i++) {
    D[i] = a.dbgr[i];
#pragma unroll
    for (int j = 0; j < P; j++) M[i][j] = a.bgcross[i * P + j];      // upper triangle is what the elimination reads
  }
  BgElim<P> e;
  bg_eliminate<P>(e, M, D);
  if (a.out_bg && g == 0 && k == 0) *a.out_bg = bg_only_logL<P>(e) + a.sls;
  if (a.G == 0) return;
  const double lp = a.lp[g];
  double* out = a.out + (size_t)g * N + k;
  if (!(lp > -INFINITY)) { *out = neg_inf(); return; }      // bayes.py:358-363: the row stays -inf
  const int h = N / 2;                                       // nsteps = int(N / 2)
  const double* m = a.tm + (size_t)g * a.NP;
  // np.correlate(x, m, 'same')[k] = sum_n x[n] m[n - k + h]
  const int n0 = max(0, k - h), n1 = min(N - 1, k - h + N - 1);
  double x[P], dm = 0.0;
#pragma unroll
  for (int j = 0; j < P; j++) x[j] = 0.0;
  for (int n = n0; n <= n1; n++) {
    const double mv = m[n - k + h];
#pragma unroll
    for (int j = 0; j < P; j++) x[j] = fma(a.abg[(size_t)j * N + n], mv, x[j]);
    dm = fma(a.dt[n], mv, dm);
  }
  // mdcross (bayes.py:367-377) with the reference's own slices: k < h: m[h-k:] on v[:len]; k >= N-h: m[:N-h-k-1] on
  // v[-len:] -- for an even N that places the template one sample late; the single middle sample of an odd N: all of it
  double mm = 0.0;
  {
    int mi0, s0, len;
    if (k < h) { mi0 = h - k; s0 = 0; len = N - mi0; }
    else if (k >= N - h) { mi0 = 0; len = 2 * N - h - k - 1; s0 = N - len; }   // the stop N-h-k-1 is negative: N + stop elements
    else { mi0 = 0; s0 = k - h; len = N; }
    for (int i = 0; i < len; i++) {
      const double mv = m[mi0 + i];
      mm = add_(mm, div_(mul_(mv, mv), a.v[s0 + i]));
    }
  }
  *out = (signal_logL<P>(e, x, mm, dm, a.half != 0) + a.sls) + lp;
}

cudaError_t launch_whole_series(int P, const double* d_tm, const double* d_abg, const double* d_dt, const double* d_v,
                                const double* d_bgcross, const double* d_dbgr, const double* d_lp, int N, int NP, int G,
                                int half, double sls, double* d_out, double* d_out_bg, cudaStream_t s, LaunchStats* st) {
  WholeArgs a{d_tm, d_abg, d_dt, d_v, d_bgcross, d_dbgr, d_lp, N, NP, G, half, sls, d_out, d_out_bg};
  const dim3 grid((unsigned)(G > 0 ? (N + 127) / 128 : 1), (unsigned)(G > 0 ? G : 1));
  switch (P) {
    case 1: k_whole_series<1><<<grid, 128, 0, s>>>(a); break;
    case 2: k_whole_series<2><<<grid, 128, 0, s>>>(a); break;
    case 3: k_whole_series<3><<<grid, 128, 0, s>>>(a); break;
    case 4: k_whole_series<4><<<grid, 128, 0, s>>>(a); break;
    case 5: k_whole_series<5><<<grid, 128, 0, s>>>(a); break;
    case 6: k_whole_series<6><<<grid, 128, 0, s>>>(a); break;
    default: return cudaErrorInvalidValue;
  }
  st->launches++;
  return cudaGetLastError();
}

struct GenArgs {
  const double* dw;     // [B][seglen]
  const double* uu;     // [B][seglen] noise variance v
  int64_t seglen, N, seg0, k0, k1;
  int B, tiles_per_curve, edges_only, tpb;
  const double* raw; const double* lp; const double* lw; const int* gorig;
  const double* bg; const double* bb;
  int G, W, WP, gchunk, NC, nhyp, sub;
  int tm_global;        // templates read from global memory (chunk too large for shared memory)
  int hyp_begin[MAXH + 1];
  int hyp_half[MAXH];
  double2* part;
  double* bgabs;        // [B*nk] or nullptr
  double* out_full;     // FULL mode
  double sls;
  int ihyp;
};

template <int P, int MODE>
__global__ void k_window_general(const GenArgs a) {
  extern __shared__ double smem[];
  const int W = a.W, H = W / 2, WP = a.WP, TPB = a.tpb;
  const int tid = threadIdx.x;
  const int tile = blockIdx.x % a.tiles_per_curve;
  const int b = blockIdx.x / a.tiles_per_curve;
  const int chunk = blockIdx.y;
  const int gb = chunk * a.gchunk, ge = min(a.G, gb + a.gchunk);
  constexpr int NPAIR = P * (P + 1) / 2;

  double* sbg = smem;                       // [P][WP]
  double* sbb = sbg + P * WP;               // [NPAIR][WP]
  double* stm = sbb + NPAIR * WP;           // [gchunk][WP]
  double* swu = stm + (size_t)(a.tm_global ? 0 : a.gchunk) * WP;   // [W][TPB]  noise-variance window, transposed
  double* swd = swu + (size_t)W * TPB;      // [W][TPB]  d/v window, transposed
  double* st64 = swd + (size_t)W * TPB;     // [64] 2^(j/64) for the log-sum-exp
  for (int i = tid; i < 64; i += TPB) st64[i] = exp2((double)i * (1.0 / 64.0));

  for (int i = tid; i < P * WP; i += TPB) sbg[i] = a.bg[i];
  for (int i = tid; i < NPAIR * WP; i += TPB) sbb[i] = a.bb[i];
  if (!a.tm_global)
    for (int i = tid; i < (ge - gb) * WP; i += TPB) stm[i] = a.raw[(size_t)gb * WP + i];
  const double* tmpl = a.tm_global ? a.raw + (size_t)gb * WP : stm;

  // which sample does this thread own?
  int64_t k;
  const int e = tile * TPB + tid;
  if (a.edges_only) {
    k = (e < H) ? e : (a.N - 2 * H + e);
    if (e >= 2 * H) k = -1;
  } else {
    k = a.k0 + e;
  }
  const bool active = (k >= a.k0 && k < a.k1 && k >= 0 && k < a.N);
  int w0 = 0, w1 = W;
  int64_t base = 0;
  if (active) {
    // window truncation at the series ends, bayes.py:320-327
    if (k < H) w0 = (int)(H - k);
    if (k >= a.N - H) w1 = (int)(W - (k + H + 1 - a.N));
    base = k - H;
    const double* du = a.uu + (size_t)b * a.seglen - a.seg0;
    const double* dd = a.dw + (size_t)b * a.seglen - a.seg0;
    for (int w = w0; w < w1; w++) {
      swu[w * TPB + tid] = du[base + w];
      swd[w * TPB + tid] = dd[base + w];
    }
  }
  __syncthreads();
  if (!active) return;
  const int n = w1 - w0;
  const int64_t nk = a.k1 - a.k0;
  const int64_t ki = (int64_t)b * nk + (k - a.k0);

  // ---- background block (bayes.py:315-329) and data x background (bayes.py:403-407) ----
  BgElim<P> el;
  {
    double M[P][P], D[P];
    int pq = 0;
#pragma unroll
    for (int p = 0; p < P; p++) {
#pragma unroll
      for (int q = p; q < P; q++) {
        const double* bbp = sbb + pq * WP + w0;
        M[p][q] = pairwise_np(n, [&](int i) { return div_(bbp[i], swu[(w0 + i) * TPB + tid]); });
        pq++;
      }
      const double* bp = sbg + p * WP + w0;
      double s = 0.0;
      for (int i = 0; i < n; i++) s = fma(swd[(w0 + i) * TPB + tid], bp[i], s);
      D[p] = s;
    }
    bg_eliminate<P>(el, M, D);
  }
  const double lnBbg = bg_only_logL<P>(el);
  if (chunk == 0 && a.bgabs) a.bgabs[ki] = lnBbg;

  const int h_lo = (MODE == 0) ? 0 : max(a.ihyp, 0);
  const int h_hi = (MODE == 0) ? a.nhyp : a.ihyp + 1;
  for (int h = h_lo; h < h_hi; h++) {
    const int lo = max(gb, a.hyp_begin[h]), hi = min(ge, a.hyp_begin[h + 1]);
    if (lo >= hi) continue;
    const bool half = a.hyp_half[h] != 0;
    double mx = LSE_EMPTY, sm = 0.0;
    for (int g = lo; g < hi; g++) {
      const double* m = tmpl + (size_t)(g - gb) * WP + w0;
      // model x model (bayes.py:367-377), model x background (bayes.py:380-393)
      const double mm = pairwise_np(n, [&](int i) { return div_(mul_(m[i], m[i]), swu[(w0 + i) * TPB + tid]); });
      double mdbg[P];
#pragma unroll
      for (int p = 0; p < P; p++) {
        const double* bp = sbg + p * WP + w0;
        mdbg[p] = pairwise_np(n, [&](int i) { return div_(mul_(bp[i], m[i]), swu[(w0 + i) * TPB + tid]); });
      }
      // data x model (general.pyx:210)
      double dm = 0.0;
      for (int i = 0; i < n; i++) dm = fma(swd[(w0 + i) * TPB + tid], m[i], dm);
      const double lnB = signal_logL<P>(el, mdbg, mm, dm, half);
      if (MODE == 0) {
        lse_push(mx, sm, (lnB - lnBbg) + (a.lp[g] + a.lw[g]), st64);
      } else {
        a.out_full[(size_t)a.gorig[g] * nk + (k - a.k0)] = (lnB + a.sls) + a.lp[g];
      }
    }
    if (MODE == 0) {
      a.part[(size_t)((h * a.NC + chunk) * a.sub) * a.B * nk + ki] = make_double2(mx, sm);
      for (int u = 1; u < a.sub; u++) a.part[(size_t)((h * a.NC + chunk) * a.sub + u) * a.B * nk + ki] = make_double2(LSE_EMPTY, 0.0);
    }
  }
}

static size_t general_smem(int P, int W, int gchunk, int tpb) {
  const int WP = W + 1;
  return sizeof(double) * ((size_t)(P + P * (P + 1) / 2 + gchunk) * WP + (size_t)2 * W * tpb + 64);
}

template <int MODE>
static cudaError_t launch_general_impl(GenArgs& a, int P, cudaStream_t s, LaunchStats* st) {
  a.tm_global = general_smem(P, a.W, a.gchunk, a.tpb) > 200 * 1024 ? 1 : 0;
  const size_t smem = general_smem(P, a.W, a.tm_global ? 0 : a.gchunk, a.tpb);
  dim3 grid((unsigned)(a.tiles_per_curve * a.B), (unsigned)a.NC);
  cudaError_t e = cudaSuccess;
#define BFL_GL(P_)                                                                                        \
  case P_:                                                                                                \
    e = cudaFuncSetAttribute(k_window_general<P_, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                             (int)smem);                                                                  \
    if (e != cudaSuccess) return e;                                                                       \
    k_window_general<P_, MODE><<<grid, a.tpb, smem, s>>>(a);                                              \
    break;
  switch (P) {
    BFL_GL(1) BFL_GL(2) BFL_GL(3) BFL_GL(4) BFL_GL(5) BFL_GL(6)
    default: return cudaErrorInvalidValue;
  }
#undef BFL_GL
  st->launches++;
  return cudaGetLastError();
}

static void fill_gen_args(GenArgs& a, const PlanView& pv, const SeriesView& sv, const Scratch& sc, bool edges_only) {
  a.dw = sc.dw; a.uu = sc.uu;
  a.seglen = sv.seglen; a.N = sv.N; a.seg0 = sv.seg0; a.k0 = sv.k0; a.k1 = sv.k1; a.B = sv.B;
  a.tpb = (pv.W <= 63) ? 64 : 32;
  a.edges_only = edges_only ? 1 : 0;
  const int64_t cnt = edges_only ? 2 * (pv.W / 2) : (sv.k1 - sv.k0);
  a.tiles_per_curve = (int)((cnt + a.tpb - 1) / a.tpb);
  if (a.tiles_per_curve < 1) a.tiles_per_curve = 1;
  a.raw = pv.raw; a.lp = pv.lp; a.lw = pv.lw; a.gorig = pv.gorig; a.bg = pv.bg; a.bb = pv.bb;
  a.G = pv.G; a.W = pv.W; a.WP = pv.WP; a.gchunk = sc.gchunk; a.NC = sc.NC; a.nhyp = pv.nhyp; a.sub = sc.sub;
  for (int h = 0; h <= pv.nhyp; h++) a.hyp_begin[h] = pv.hyp_begin[h];
  for (int h = 0; h < pv.nhyp; h++) a.hyp_half[h] = pv.hyp_half[h];
  a.part = sc.part; a.bgabs = sc.bgabs; a.out_full = nullptr; a.sls = 0.0; a.ihyp = 0;
}

cudaError_t launch_window_general_lse(const PlanView& pv, const SeriesView& sv, Scratch& sc, bool edges_only,
                                      cudaStream_t s, LaunchStats* st) {
  GenArgs a;
  fill_gen_args(a, pv, sv, sc, edges_only);
  return launch_general_impl<0>(a, pv.P, s, st);
}

cudaError_t launch_window_general_full(const PlanView& pv, int ihyp, const SeriesView& sv, Scratch& sc,
                                       bool edges_only, double sumlogsigma, double* d_out,
                                       cudaStream_t s, LaunchStats* st) {
  GenArgs a;
  fill_gen_args(a, pv, sv, sc, edges_only);
  a.out_full = d_out; a.sls = sumlogsigma; a.ihyp = ihyp;
  return launch_general_impl<1>(a, pv.P, s, st);
}

// ======================================================================================
// WEIGHTED path: interior samples of a curve with per-sample variances v_j = sk^2 + cle_j^2
// ======================================================================================
// Same integral as the general path (bayes.py:315-407 -> general.pyx:143-246 -> log_marg_amp_full_C),
// evaluated in the orthonormal polynomial basis Q of the plan, where the system is well conditioned
// (full windows only -- the truncated edge windows keep the reference-order arithmetic above):
//   A = Q U Q^T (P x P), c = Q U d, per template e = m^T U m, f = Q U m, r0 = m^T U d, U = diag(1/v)
//   A = L L^T, yc = L^-1 c, yf = L^-1 f, s = e - |yf|^2, r = r0 - yf.yc, z = r / sqrt(s)
//   lnB_bg = -0.5 logdet(G) - sum log L_ii + 0.5 |yc|^2 + 0.5 P log 2 pi
//   lnB_g - lnB_bg = -0.5 log s + { 0.5 log(pi/2) + phi(z) | 0.5 log(2 pi) + z^2/2 }
// One thread per sample; the 1/v and d/v tiles (+ halo) live in shared memory and are read with
// lane-consecutive (conflict-free) loads, Q and the templates with broadcast 128-bit loads; WGT_GT
// templates share each pass over the window, so a pass is (P+2) WGT_GT FMAs per tap against
// (2 + (P + WGT_GT)/2) shared-memory loads.  Algorithmic work: (P + 2) W FMAs per (sample, template).
struct WgtArgs {
  const double* dw;     // [B][seglen]  d / v
  const double* uu;     // [B][seglen]  v
  int64_t seglen, N, seg0, k0, k1;
  int B, tiles_per_curve;
  const double* raw; const double* qb; const double* lp; const double* lw; const int* gorig;
  int G, W, WP, gchunk, NC, nhyp, sub;
  int hyp_begin[MAXH + 1];
  int hyp_half[MAXH];
  double logdetG;
  double2* part;
  double* bgabs;
  double* out_full;
  double sls;
  int ihyp;
};

constexpr int WGT_TPB = 128, WGT_GT = 4;

template <int P, int MODE>
__global__ void __launch_bounds__(WGT_TPB) k_window_weighted(const WgtArgs a) {
  extern __shared__ double smem[];
  const int W = a.W, H = W / 2, WP = a.WP;
  const int tid = threadIdx.x;
  const int tile = blockIdx.x % a.tiles_per_curve, b = blockIdx.x / a.tiles_per_curve;
  const int chunk = blockIdx.y;
  const int gb = chunk * a.gchunk, ge = min(a.G, gb + a.gchunk);
  const int nd = (WGT_TPB + 2 * H + 2) & ~1;
  double* su = smem;                               // [nd] 1/v
  double* sud = su + nd;                           // [nd] d/v
  double* sq = sud + nd;                           // [P][WP]
  double* stm = sq + P * WP;                       // [gchunk][WP]
  double* sphi = stm + (size_t)a.gchunk * WP;      // [(DEG+1) * NI]
  double* st64 = sphi + BFL_PHI_CM_SIZE;           // [64]

  const int64_t kbase = a.k0 + (int64_t)tile * WGT_TPB;
  {
    const double* du = a.uu + (size_t)b * a.seglen;
    const double* dd = a.dw + (size_t)b * a.seglen;
    const int64_t lo = a.seg0, hi = min(a.N, a.seg0 + a.seglen);
    for (int i = tid; i < WGT_TPB + 2 * H; i += WGT_TPB) {
      const int64_t sidx = kbase - H + i;
      const bool in = sidx >= lo && sidx < hi;
      su[i] = in ? 1.0 / du[sidx - a.seg0] : 0.0;
      sud[i] = in ? dd[sidx - a.seg0] : 0.0;
    }
    for (int i = tid; i < P * WP; i += WGT_TPB) sq[i] = a.qb[i];
    for (int i = tid; i < (ge - gb) * WP; i += WGT_TPB) stm[i] = a.raw[(size_t)gb * WP + i];
    if (tid < 64) st64[tid] = exp2((double)tid * (1.0 / 64.0));
    for (int i = tid; i < BFL_PHI_CM_SIZE; i += WGT_TPB) {
      const int cidx = i / BFL_PHI_NI, kk = i - cidx * BFL_PHI_NI;
      sphi[i] = BFL_PHI_TAB[kk * BFL_PHI_ROW + cidx];
    }
  }
  __syncthreads();
  const int64_t k = kbase + tid;
  if (k >= a.k1 || k < H || k >= a.N - H) return;      // interior samples only
  const int64_t nk = a.k1 - a.k0;
  const int64_t ki = (int64_t)b * nk + (k - a.k0);
  const double* uw = su + tid;
  const double* dw = sud + tid;

  // ---- shared by every template of this sample: A = Q U Q^T, c = Q U d ----
  double L[P][P], inv[P], yc[P];
  double logdetA = 0.0, quad = 0.0;
  {
    double A[P][P], c[P];
#pragma unroll
    for (int p = 0; p < P; p++) {
      c[p] = 0.0;
#pragma unroll
      for (int q = 0; q < P; q++) A[p][q] = 0.0;
    }
    for (int j = 0; j < W; j++) {
      const double u = uw[j], ud = dw[j];
      double qv[P];
#pragma unroll
      for (int p = 0; p < P; p++) qv[p] = sq[p * WP + j];
#pragma unroll
      for (int p = 0; p < P; p++) {
        const double uq = u * qv[p];
        c[p] = fma(ud, qv[p], c[p]);
#pragma unroll
        for (int q = p; q < P; q++) A[p][q] = fma(uq, qv[q], A[p][q]);
      }
    }
#pragma unroll
    for (int i = 0; i < P; i++) {
      double d = A[i][i];
#pragma unroll
      for (int t = 0; t < i; t++) d = fma(-L[i][t], L[i][t], d);
      logdetA += log(d);
      const double r = rsqrt(d);
      inv[i] = r;
      L[i][i] = d * r;
#pragma unroll
      for (int rr = i + 1; rr < P; rr++) {
        double x = A[i][rr];
#pragma unroll
        for (int t = 0; t < i; t++) x = fma(-L[rr][t], L[i][t], x);
        L[rr][i] = x * r;
      }
      double y = c[i];
#pragma unroll
      for (int t = 0; t < i; t++) y = fma(-L[i][t], yc[t], y);
      yc[i] = y * r;
      quad = fma(yc[i], yc[i], quad);
    }
  }
  double lnBbg = -0.5 * a.logdetG - 0.5 * logdetA + 0.5 * quad + 0.5 * P * (BFL_LN2 + BFL_LNPI);
  if (!isfinite(lnBbg)) lnBbg = neg_inf();          // non-finite data or variances, log_marg_amp_full.c:20,25
  if (chunk == 0 && a.bgabs) a.bgabs[ki] = lnBbg;

  const int h_lo = (MODE == 0) ? 0 : max(a.ihyp, 0);
  const int h_hi = (MODE == 0) ? a.nhyp : a.ihyp + 1;
  const int NP2 = W / 2;
  for (int h = h_lo; h < h_hi; h++) {
    const int lo = max(gb, a.hyp_begin[h]), hi = min(ge, a.hyp_begin[h + 1]);
    if (lo >= hi) continue;
    const bool half = a.hyp_half[h] != 0;
    const double hc = half ? 0.5 * (BFL_LNPI - BFL_LN2) : 0.5 * (BFL_LN2 + BFL_LNPI);
    double mx = LSE_EMPTY, sm = 0.0;
    for (int g4 = lo; g4 < hi; g4 += WGT_GT) {
      const double* mrow[WGT_GT];
#pragma unroll
      for (int t = 0; t < WGT_GT; t++) mrow[t] = stm + (size_t)(min(g4 + t, hi - 1) - gb) * WP;
      double e[WGT_GT], r0[WGT_GT], f[WGT_GT][P];
#pragma unroll
      for (int t = 0; t < WGT_GT; t++) {
        e[t] = 0.0; r0[t] = 0.0;
#pragma unroll
        for (int p = 0; p < P; p++) f[t][p] = 0.0;
      }
      for (int j2 = 0; j2 < NP2; j2++) {
        const double u0 = uw[2 * j2], u1 = uw[2 * j2 + 1], d0 = dw[2 * j2], d1 = dw[2 * j2 + 1];
        double2 qv[P];
#pragma unroll
        for (int p = 0; p < P; p++) qv[p] = reinterpret_cast<const double2*>(sq + p * WP)[j2];
#pragma unroll
        for (int t = 0; t < WGT_GT; t++) {
          const double2 m = reinterpret_cast<const double2*>(mrow[t])[j2];
          const double w0 = u0 * m.x, w1 = u1 * m.y;
          e[t] = fma(w0, m.x, e[t]); e[t] = fma(w1, m.y, e[t]);
          r0[t] = fma(d0, m.x, r0[t]); r0[t] = fma(d1, m.y, r0[t]);
#pragma unroll
          for (int p = 0; p < P; p++) { f[t][p] = fma(w0, qv[p].x, f[t][p]); f[t][p] = fma(w1, qv[p].y, f[t][p]); }
        }
      }
      {
        const int j = W - 1;
        const double u0 = uw[j], d0 = dw[j];
#pragma unroll
        for (int t = 0; t < WGT_GT; t++) {
          const double m = mrow[t][j];
          const double w0 = u0 * m;
          e[t] = fma(w0, m, e[t]);
          r0[t] = fma(d0, m, r0[t]);
#pragma unroll
          for (int p = 0; p < P; p++) f[t][p] = fma(w0, sq[p * WP + j], f[t][p]);
        }
      }
#pragma unroll
      for (int t = 0; t < WGT_GT; t++) {
        const int g = g4 + t;
        if (g >= hi) break;
        double s = e[t], r = r0[t];
        double yf[P];
#pragma unroll
        for (int i = 0; i < P; i++) {
          double y = f[t][i];
#pragma unroll
          for (int q = 0; q < i; q++) y = fma(-L[i][q], yf[q], y);
          yf[i] = y * inv[i];
          s = fma(-yf[i], yf[i], s);
          r = fma(-yf[i], yc[i], r);
        }
        const double z = r * rsqrt(s);
        double val;
        if (half) {
          val = phi_tab(z, sphi);
          if (z < BFL_PHI_ZFAR) val = phi_half_far(z);
        } else {
          val = 0.5 * z * z;
        }
        val += hc - 0.5 * log(s);
        if (lnBbg == neg_inf()) val = lnBbg - lnBbg;      // NaN, as (-inf) - (-inf) in the general path
        if (MODE == 0) {
          lse_push(mx, sm, val + (a.lp[g] + a.lw[g]), st64);
        } else {
          a.out_full[(size_t)a.gorig[g] * nk + (k - a.k0)] =
              (lnBbg == neg_inf()) ? lnBbg : ((val + lnBbg) + a.sls) + a.lp[g];
        }
      }
    }
    if (MODE == 0) {
      a.part[(size_t)((h * a.NC + chunk) * a.sub) * a.B * nk + ki] = make_double2(mx, sm);
      for (int u = 1; u < a.sub; u++)
        a.part[(size_t)((h * a.NC + chunk) * a.sub + u) * a.B * nk + ki] = make_double2(LSE_EMPTY, 0.0);
    }
  }
}

static size_t weighted_smem(int P, int W, int gchunk) {
  const int WP = W + 1;
  const int nd = (WGT_TPB + 2 * (W / 2) + 2) & ~1;
  return sizeof(double) * ((size_t)2 * nd + (size_t)(P + gchunk) * WP + BFL_PHI_CM_SIZE + 64);
}

bool weighted_fits(const PlanView& pv, const Scratch& sc) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("BFL_NO_WEIGHTED"); off = (e && atoi(e)) ? 1 : 0; }
  return !off && pv.P >= 1 && pv.P <= 6 && (pv.WP % 2) == 0 && weighted_smem(pv.P, pv.W, sc.gchunk) <= 220 * 1024;
}

template <int MODE>
static cudaError_t launch_weighted_impl(WgtArgs& a, int P, cudaStream_t s, LaunchStats* st) {
  const size_t smem = weighted_smem(P, a.W, a.gchunk);
  dim3 grid((unsigned)(a.tiles_per_curve * a.B), (unsigned)a.NC);
  cudaError_t e = cudaSuccess;
#define BFL_WL(P_)                                                                                        \
  case P_:                                                                                                \
    e = cudaFuncSetAttribute(k_window_weighted<P_, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                             (int)smem);                                                                  \
    if (e != cudaSuccess) return e;                                                                       \
    k_window_weighted<P_, MODE><<<grid, WGT_TPB, smem, s>>>(a);                                           \
    break;
  switch (P) {
    BFL_WL(1) BFL_WL(2) BFL_WL(3) BFL_WL(4) BFL_WL(5) BFL_WL(6)
    default: return cudaErrorInvalidValue;
  }
#undef BFL_WL
  st->launches++;
  return cudaGetLastError();
}

static void fill_wgt_args(WgtArgs& a, const PlanView& pv, const SeriesView& sv, const Scratch& sc) {
  a.dw = sc.dw; a.uu = sc.uu;
  a.seglen = sv.seglen; a.N = sv.N; a.seg0 = sv.seg0; a.k0 = sv.k0; a.k1 = sv.k1; a.B = sv.B;
  a.tiles_per_curve = (int)((sv.k1 - sv.k0 + WGT_TPB - 1) / WGT_TPB);
  a.raw = pv.raw; a.qb = pv.qb; a.lp = pv.lp; a.lw = pv.lw; a.gorig = pv.gorig;
  a.G = pv.G; a.W = pv.W; a.WP = pv.WP; a.gchunk = sc.gchunk; a.NC = sc.NC; a.nhyp = pv.nhyp; a.sub = sc.sub;
  for (int h = 0; h <= pv.nhyp; h++) a.hyp_begin[h] = pv.hyp_begin[h];
  for (int h = 0; h < pv.nhyp; h++) a.hyp_half[h] = pv.hyp_half[h];
  a.logdetG = pv.logdetG;
  a.part = sc.part; a.bgabs = sc.bgabs; a.out_full = nullptr; a.sls = 0.0; a.ihyp = 0;
}

cudaError_t launch_window_weighted_lse(const PlanView& pv, const SeriesView& sv, Scratch& sc, cudaStream_t s,
                                       LaunchStats* st) {
  WgtArgs a;
  fill_wgt_args(a, pv, sv, sc);
  return launch_weighted_impl<0>(a, pv.P, s, st);
}

cudaError_t launch_window_weighted_full(const PlanView& pv, int ihyp, const SeriesView& sv, Scratch& sc,
                                        double sumlogsigma, double* d_out, cudaStream_t s, LaunchStats* st) {
  WgtArgs a;
  fill_wgt_args(a, pv, sv, sc);
  a.out_full = d_out; a.sls = sumlogsigma; a.ihyp = ihyp;
  return launch_weighted_impl<1>(a, pv.P, s, st);
}

// --------------------------------------------------------------------------------------
// WEIGHTED path, separable form (PlanView::rshapes).  A Flare template is rise (+) centre (+) decay on
// disjoint supports, so with m_g = A_i + B_j every weighted sum separates -- including the norm:
//   e_g = e_Ai + e_Bj,  f_g = f_Ai + f_Bj,  r0_g = r0_Ai + r0_Bj
// and the triangular solve is linear, so per SHAPE (not per template) the kernel forms
//   y = L^-1 f,  s' = e - |y|^2,  r' = r0 - y.yc
// leaving s_g = s'_A + s'_B - 2 yA.yB, r_g = r'_A + r'_B per template.  The -0.5 log s_g term rides
// along as a weight 1/sqrt(s_g) in the log-sum-exp, so a template costs P+4 flops, one rsqrt, the
// phi look-up and one exp.  Shapes are summed over their support only (27 taps for a rise, 28 for
// a decay, 1 for an impulse).  One thread per sample; A-shape results live in a per-thread
// shared-memory slice, B shapes are handled WSEP_GT at a time in registers.

constexpr int WSEP_TPB = 256, WSEP_GT = 4;

// Persistent CTAs (one per SM): tables staged once, then a static stride over tiles of WSEP_TPB samples.
template <int P>
__global__ void __launch_bounds__(WSEP_TPB) k_window_wsep(const WsepArgs a) {
  extern __shared__ double smem[];
  constexpr int NV = P + 2;                        // per shape: y[P], s', r'
  const int W = a.W, H = W / 2, WP = a.WP;
  const int tid = threadIdx.x;
  const int nd = (WSEP_TPB + 2 * H + 2) & ~1;
  double* su = smem;                               // [nd] 1/v
  double* sud = su + nd;                           // [nd] d/v
  double* sq = sud + nd;                           // [P][WP]
  double* ssh = sq + P * WP;                       // [S][WP]
  double* sK = ssh + (size_t)a.S * WP;             // [G] lp + lw
  double* sphi = sK + ((a.G + 1) & ~1);            // [(DEG+1) * NI]
  double* st64 = sphi + BFL_PHI_CM_SIZE;           // [64]
  double* ysm = st64 + 64;                         // [arows][NV][TPB]
  int* sG01 = reinterpret_cast<int*>(ysm + (size_t)a.arows * NV * WSEP_TPB);   // [S][2] templates using shape s as B
  short* sIdxA = reinterpret_cast<short*>(sG01 + 2 * a.S);                     // [G]
  short* sLoHi = sIdxA + ((a.G + 7) & ~7);                                     // [S][2]

  for (int i = tid; i < P * WP; i += WSEP_TPB) sq[i] = a.qb[i];
  for (int i = tid; i < a.S * WP; i += WSEP_TPB) ssh[i] = a.rshapes[i];
  for (int i = tid; i < a.G; i += WSEP_TPB) { sK[i] = a.lp[i] + a.lw[i]; sIdxA[i] = a.idxA[i]; }
  for (int i = tid; i < a.S; i += WSEP_TPB) {
    sLoHi[2 * i] = a.sh_lo[i]; sLoHi[2 * i + 1] = a.sh_hi[i];
    sG01[2 * i] = a.sh_g0[i]; sG01[2 * i + 1] = a.sh_g1[i];
  }
  if (tid < 64) st64[tid] = exp2((double)tid * (1.0 / 64.0));
  for (int i = tid; i < BFL_PHI_CM_SIZE; i += WSEP_TPB) {
    const int cidx = i / BFL_PHI_NI, kk = i - cidx * BFL_PHI_NI;
    sphi[i] = BFL_PHI_TAB[kk * BFL_PHI_ROW + cidx];
  }
  const int64_t nk = a.k1 - a.k0;
  const double* uw = su + tid;
  const double* dw = sud + tid;

  for (int64_t tl = blockIdx.x; tl < a.total_tiles; tl += gridDim.x) {
  const int tile = (int)(tl % a.tiles_per_curve), b = (int)(tl / a.tiles_per_curve);
  const int64_t kbase = a.k0 + (int64_t)tile * WSEP_TPB;
  __syncthreads();                                 // the previous tile is done with su / sud
  {
    const double* du = a.uu + (size_t)b * a.seglen;
    const double* dd = a.dw + (size_t)b * a.seglen;
    const int64_t lo = a.seg0, hi = min(a.N, a.seg0 + a.seglen);
    for (int i = tid; i < nd; i += WSEP_TPB) {
      const int64_t sidx = kbase - H + i;
      const bool in = i < WSEP_TPB + 2 * H && sidx >= lo && sidx < hi;
      su[i] = in ? 1.0 / du[sidx - a.seg0] : 0.0;
      sud[i] = in ? dd[sidx - a.seg0] : 0.0;
    }
  }
  __syncthreads();
  const int64_t k = kbase + tid;
  if (k >= a.k1 || k < H || k >= a.N - H) continue;    // interior samples only (no barrier below this line)
  const int64_t ki = (int64_t)b * nk + (k - a.k0);

  // ---- A = Q U Q^T = L L^T, yc = L^-1 Q U d ----
  double L[P][P], inv[P], yc[P];
  double logdetA = 0.0, quad = 0.0;
  {
    double A[P][P], c[P];
#pragma unroll
    for (int p = 0; p < P; p++) {
      c[p] = 0.0;
#pragma unroll
      for (int q = 0; q < P; q++) A[p][q] = 0.0;
    }
    for (int j = 0; j < W; j++) {
      const double u = uw[j], ud = dw[j];
      double qv[P];
#pragma unroll
      for (int p = 0; p < P; p++) qv[p] = sq[p * WP + j];
#pragma unroll
      for (int p = 0; p < P; p++) {
        const double uq = u * qv[p];
        c[p] = fma(ud, qv[p], c[p]);
#pragma unroll
        for (int q = p; q < P; q++) A[p][q] = fma(uq, qv[q], A[p][q]);
      }
    }
#pragma unroll
    for (int i = 0; i < P; i++) {
      double d = A[i][i];
#pragma unroll
      for (int t = 0; t < i; t++) d = fma(-L[i][t], L[i][t], d);
      logdetA += log(d);
      const double r = rsqrt(d);
      inv[i] = r;
      L[i][i] = d * r;
#pragma unroll
      for (int rr = i + 1; rr < P; rr++) {
        double x = A[i][rr];
#pragma unroll
        for (int t = 0; t < i; t++) x = fma(-L[rr][t], L[i][t], x);
        L[rr][i] = x * r;
      }
      double y = c[i];
#pragma unroll
      for (int t = 0; t < i; t++) y = fma(-L[i][t], yc[t], y);
      yc[i] = y * r;
      quad = fma(yc[i], yc[i], quad);
    }
  }
  double lnBbg = -0.5 * a.logdetG - 0.5 * logdetA + 0.5 * quad + 0.5 * P * (BFL_LN2 + BFL_LNPI);
  const bool bad = !isfinite(lnBbg);               // non-finite data or variances, log_marg_amp_full.c:20,25
  if (bad) lnBbg = neg_inf();
  if (a.bgabs) a.bgabs[ki] = lnBbg;

  // weighted sums of WSEP_GT shapes over the union of their supports, reduced to (y, s', r')
  auto shape_group = [&](int s0, int smax, double (&V)[WSEP_GT][NV]) {
    const double* row[WSEP_GT];
    int lo = WP, hi = 0;
#pragma unroll
    for (int t = 0; t < WSEP_GT; t++) {
      const int si = min(s0 + t, smax);
      row[t] = ssh + (size_t)si * WP;
      lo = min(lo, (int)sLoHi[2 * si]);
      hi = max(hi, (int)sLoHi[2 * si + 1]);
    }
    double e[WSEP_GT], r0[WSEP_GT], f[WSEP_GT][P];
#pragma unroll
    for (int t = 0; t < WSEP_GT; t++) {
      e[t] = 0.0; r0[t] = 0.0;
#pragma unroll
      for (int p = 0; p < P; p++) f[t][p] = 0.0;
    }
    for (int j2 = lo >> 1; j2 < (hi >> 1); j2++) {
      const double u0 = uw[2 * j2], u1 = uw[2 * j2 + 1], d0 = dw[2 * j2], d1 = dw[2 * j2 + 1];
      double2 qv[P];
#pragma unroll
      for (int p = 0; p < P; p++) qv[p] = reinterpret_cast<const double2*>(sq + p * WP)[j2];
#pragma unroll
      for (int t = 0; t < WSEP_GT; t++) {
        const double2 m = reinterpret_cast<const double2*>(row[t])[j2];
        const double w0 = u0 * m.x, w1 = u1 * m.y;
        e[t] = fma(w0, m.x, e[t]); e[t] = fma(w1, m.y, e[t]);
        r0[t] = fma(d0, m.x, r0[t]); r0[t] = fma(d1, m.y, r0[t]);
#pragma unroll
        for (int p = 0; p < P; p++) { f[t][p] = fma(w0, qv[p].x, f[t][p]); f[t][p] = fma(w1, qv[p].y, f[t][p]); }
      }
    }
    if (hi & 1) {   // odd end: one more tap (never reads past the window, so a NaN beyond it stays out)
      const int j = hi - 1;
      const double u0 = uw[j], d0 = dw[j];
#pragma unroll
      for (int t = 0; t < WSEP_GT; t++) {
        const double m = row[t][j];
        const double w0 = u0 * m;
        e[t] = fma(w0, m, e[t]);
        r0[t] = fma(d0, m, r0[t]);
#pragma unroll
        for (int p = 0; p < P; p++) f[t][p] = fma(w0, sq[p * WP + j], f[t][p]);
      }
    }
#pragma unroll
    for (int t = 0; t < WSEP_GT; t++) {
      double s = e[t], r = r0[t];
#pragma unroll
      for (int i = 0; i < P; i++) {
        double y = f[t][i];
#pragma unroll
        for (int q = 0; q < i; q++) y = fma(-L[i][q], V[t][q], y);
        y *= inv[i];
        V[t][i] = y;
        s = fma(-y, y, s);
        r = fma(-y, yc[i], r);
      }
      V[t][P] = s;
      V[t][P + 1] = r;
    }
  };

  for (int h = 0; h < a.nhyp; h++) {
    if (a.hyp_begin[h] >= a.hyp_begin[h + 1]) continue;
    const bool half = a.hyp_half[h] != 0;
    const double hc = half ? 0.5 * (BFL_LNPI - BFL_LN2) : 0.5 * (BFL_LN2 + BFL_LNPI);
    const int nA = a.hyp_nA[h], A0 = a.hyp_A0[h], B0 = a.hyp_B0[h], nB = a.hyp_nB[h];
    double mx = LSE_EMPTY, sm = 0.0;
    // WSEP_GT templates at a time: independent chains through rsqrt, phi and exp, one merge into (mx, sm)
    auto term = [&](double s, double r, double kg, bool valid, double& x, double& w) {
      w = rsqrt(s);
      const double z = r * w;
      double val;
      if (half) {
        val = phi_tab(z, sphi);
        if (z < BFL_PHI_ZFAR) val = phi_half_far(z);
      } else {
        val = 0.5 * z * z;
      }
      x = valid ? val + kg : LSE_EMPTY;
      w = valid ? w : 0.0;
    };
    auto merge = [&](const double (&x)[WSEP_GT], const double (&w)[WSEP_GT]) {
      double bm = x[0];
#pragma unroll
      for (int c = 1; c < WSEP_GT; c++) bm = fmax(bm, x[c]);
      double bs = 0.0;
#pragma unroll
      for (int c = 0; c < WSEP_GT; c++) bs = fma(w[c], exp_nonpos(x[c] - bm, st64), bs);
      const double d = bm - mx;
      const double e = exp_nonpos(-fabs(d), st64);
      const bool up = __double2hiint(d) >= 0;
      sm = up ? fma(sm, e, bs) : fma(bs, e, sm);
      mx = up ? bm : mx;
    };
    const int npass = nA > 0 ? (nA + a.arows - 1) / a.arows : 1;
    for (int pass = 0; pass < npass; pass++) {
      const int ia0 = pass * a.arows, ia1 = min(nA, ia0 + a.arows);
      for (int ia = ia0; ia < ia1; ia += WSEP_GT) {
        double V[WSEP_GT][NV];
        shape_group(A0 + ia, A0 + ia1 - 1, V);
#pragma unroll
        for (int t = 0; t < WSEP_GT; t++) {
          if (ia + t < ia1) {
#pragma unroll
            for (int q = 0; q < NV; q++) ysm[((size_t)(ia + t - ia0) * NV + q) * WSEP_TPB + tid] = V[t][q];
          }
        }
      }
      for (int ib = 0; ib < nB; ib += WSEP_GT) {
        double V[WSEP_GT][NV];
        shape_group(B0 + ib, B0 + nB - 1, V);
        if (nA > 0) {
          // separable hypothesis: every B shape pairs with a run of templates (one per admissible A)
#pragma unroll
          for (int t = 0; t < WSEP_GT; t++) {
            if (ib + t >= nB) break;
            const int g0 = sG01[2 * (B0 + ib + t)], g1 = sG01[2 * (B0 + ib + t) + 1];
            for (int g = g0; g < g1; g += WSEP_GT) {
              double x[WSEP_GT], w[WSEP_GT];
#pragma unroll
              for (int c = 0; c < WSEP_GT; c++) {
                const int gi = min(g + c, g1 - 1);
                const int ia = sIdxA[gi];
                const bool valid = (g + c < g1) && ia >= ia0 && ia < ia1;
                const double* ya = ysm + (size_t)min(max(ia - ia0, 0), a.arows - 1) * NV * WSEP_TPB + tid;
                double dot = 0.0;
#pragma unroll
                for (int q = 0; q < P; q++) dot = fma(ya[q * WSEP_TPB], V[t][q], dot);
                const double s = (V[t][P] + ya[P * WSEP_TPB]) - 2.0 * dot;
                const double r = V[t][P + 1] + ya[(P + 1) * WSEP_TPB];
                term(s, r, sK[gi], valid, x[c], w[c]);
              }
              merge(x, w);
            }
          }
        } else {
          // dense hypothesis: one template per shape, the group is the batch
          double x[WSEP_GT], w[WSEP_GT];
#pragma unroll
          for (int t = 0; t < WSEP_GT; t++) {
            const int sb = min(B0 + ib + t, B0 + nB - 1);
            const int gi = sG01[2 * sb];
            term(V[t][P], V[t][P + 1], sK[gi], ib + t < nB && sG01[2 * sb + 1] > gi, x[t], w[t]);
          }
          merge(x, w);
        }
      }
    }
    if (bad) { mx = lnBbg - lnBbg; sm = mx; }      // NaN, as (-inf) - (-inf) in the general path
    a.part[(size_t)h * a.B * nk + ki] = make_double2(mx + hc, sm);
  }
  }   // tile loop
}

static size_t wsep_smem(int P, int W, int G, int S, int arows) {
  const int WP = W + 1;
  const int nd = (WSEP_TPB + 2 * (W / 2) + 2) & ~1;
  return sizeof(double) * ((size_t)2 * nd + (size_t)(P + S) * WP + ((G + 1) & ~1) + BFL_PHI_CM_SIZE + 64 +
                           (size_t)arows * (P + 2) * WSEP_TPB) +
         sizeof(int) * (size_t)(2 * S) + sizeof(short) * (size_t)(((G + 7) & ~7) + 2 * S + 8);
}

// A rows per pass: as many as fit in shared memory, at most the largest hypothesis needs
static int wsep_arows(const PlanView& pv) {
  int need = 1;
  for (int h = 0; h < pv.nhyp; h++) need = std::max(need, pv.hyp_nA[h]);
  const size_t budget = 220 * 1024;
  const size_t fixed = wsep_smem(pv.P, pv.W, pv.G, pv.S, 0);
  if (fixed >= budget) return 0;
  int fit = (int)((budget - fixed) / (sizeof(double) * (size_t)(pv.P + 2) * WSEP_TPB));
  fit &= ~(WSEP_GT - 1);
  return std::min((need + WSEP_GT - 1) & ~(WSEP_GT - 1), fit);
}

bool wsep_fits(const PlanView& pv) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("BFL_NO_WSEP"); off = (e && atoi(e)) ? 1 : 0; }
  return !off && pv.wsep_ok && pv.W <= 55 && pv.G > 0 && pv.P >= 1 && pv.P <= 6 && wsep_arows(pv) >= WSEP_GT;
}

cudaError_t launch_window_wsep_lse(const PlanView& pv, const SeriesView& sv, Scratch& sc, cudaStream_t s,
                                   LaunchStats* st) {
  WsepArgs a;
  a.dw = sc.dw; a.uu = sc.uu;
  a.seglen = sv.seglen; a.N = sv.N; a.seg0 = sv.seg0; a.k0 = sv.k0; a.k1 = sv.k1; a.B = sv.B;
  a.tiles_per_curve = (int)((sv.k1 - sv.k0 + WSEP_TPB - 1) / WSEP_TPB);
  a.total_tiles = (int64_t)a.tiles_per_curve * sv.B;
  a.rshapes = pv.rshapes; a.sh_lo = pv.sh_lo; a.sh_hi = pv.sh_hi; a.sh_g0 = pv.sh_g0; a.sh_g1 = pv.sh_g1;
  a.qb = pv.qb; a.lp = pv.lp; a.lw = pv.lw; a.idxA = pv.idxA;
  a.G = pv.G; a.W = pv.W; a.WP = pv.WP; a.S = pv.S; a.nhyp = pv.nhyp; a.arows = wsep_arows(pv);
  for (int h = 0; h <= pv.nhyp; h++) a.hyp_begin[h] = pv.hyp_begin[h];
  for (int h = 0; h < pv.nhyp; h++) {
    a.hyp_half[h] = pv.hyp_half[h];
    a.hyp_A0[h] = pv.hyp_A0[h]; a.hyp_nA[h] = pv.hyp_nA[h]; a.hyp_B0[h] = pv.hyp_B0[h]; a.hyp_nB[h] = pv.hyp_nB[h];
  }
  a.logdetG = pv.logdetG;
  a.part = sc.part; a.bgabs = sc.bgabs;
  const size_t smem = wsep_smem(pv.P, pv.W, pv.G, pv.S, a.arows);
  const int sms = sm_count_current();
  cudaError_t e = cudaSuccess;
  // large plans of the default window length: split form (bfl_wsplit.cu), Yw sized by setup_scratch
  if (sc.Yw) return launch_window_wsplit(pv, a, sc.Yw, sc.Yw_curves, sms, sc.Yw_lnO, sc.Yw_noisepoly, sc.split_counters, s, st);
  dim3 grid((unsigned)std::min<int64_t>(a.total_tiles, sms));   // persistent: one CTA per SM (shared-memory bound)
#define BFL_WS(P_)                                                                                              \
  case P_:                                                                                                      \
    e = cudaFuncSetAttribute(k_window_wsep<P_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
    if (e != cudaSuccess) return e;                                                                             \
    k_window_wsep<P_><<<grid, WSEP_TPB, smem, s>>>(a);                                                          \
    break;
  switch (pv.P) {
    BFL_WS(1) BFL_WS(2) BFL_WS(3) BFL_WS(4) BFL_WS(5) BFL_WS(6)
    default: return cudaErrorInvalidValue;
  }
#undef BFL_WS
  st->launches++;
  return cudaGetLastError();
}

// ======================================================================================
// combine: merge chunk partials -> marginal per hypothesis -> log odds (find.py:853-862)
// ======================================================================================
struct CombArgs {
  const double2* part;
  const double* bgabs;
  int64_t BK;          // B * nk
  int nhyp, NC, gchunk, G, noisepoly, sub;   // sub: partial-sum slots per template chunk (1 for every current kernel)
  int hyp_begin[MAXH + 1];
  double* lnO;
  double* marg;        // [(nhyp+1)][BK] or nullptr
};

__global__ void k_combine(const CombArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.BK) return;
  double rel[MAXH];
  for (int h = 0; h < a.nhyp; h++) {
    double M = neg_inf();
    const int c0 = a.hyp_begin[h] / a.gchunk;
    const int c1 = (a.hyp_begin[h + 1] > a.hyp_begin[h]) ? (a.hyp_begin[h + 1] - 1) / a.gchunk : c0 - 1;
    for (int c = c0 * a.sub; c < (c1 + 1) * a.sub; c++) {
      const double2 p = a.part[(size_t)(h * a.NC * a.sub + c) * a.BK + i];
      if (p.y > 0.0 || isnan(p.y)) M = fmax(M, p.x);
    }
    double S = 0.0;
    bool nan_ = false;
    for (int c = c0 * a.sub; c < (c1 + 1) * a.sub; c++) {
      const double2 p = a.part[(size_t)(h * a.NC * a.sub + c) * a.BK + i];
      if (isnan(p.x) || isnan(p.y)) nan_ = true;
      if (p.y > 0.0) S += p.y * exp(p.x - M);
    }
    rel[h] = nan_ ? nan("") : ((S > 0.0) ? M + log(S) : neg_inf());
  }
  // denominator: logplus fold in the reference's order [poly, noise models...], start from -inf
  double lnO = rel[0];
  const int nnoise = a.nhyp - 1 + (a.noisepoly ? 1 : 0);
  if (nnoise > 0) {
    double denom = neg_inf();
    if (a.noisepoly) denom = logplus_dev(denom, 0.0);
    for (int h = 1; h < a.nhyp; h++) denom = logplus_dev(denom, rel[h]);
    lnO = rel[0] - denom;
  } else if (a.bgabs) {
    lnO = rel[0] + a.bgabs[i];   // no noise models: signal versus Gaussian noise (find.py:859-862)
  }
  a.lnO[i] = lnO;
  if (a.marg) {
    const double bgv = a.bgabs ? a.bgabs[i] : 0.0;
    for (int h = 0; h < a.nhyp; h++) a.marg[(size_t)h * a.BK + i] = rel[h] + bgv;
    a.marg[(size_t)a.nhyp * a.BK + i] = bgv;
  }
}

cudaError_t launch_combine(const PlanView& pv, const SeriesView& sv, const Scratch& sc, int noisepoly,
                           double* d_lnO, double* d_marg, cudaStream_t s, LaunchStats* st) {
  CombArgs a;
  a.part = sc.part; a.bgabs = sc.bgabs;
  a.BK = (int64_t)sv.B * (sv.k1 - sv.k0);
  a.nhyp = pv.nhyp; a.NC = sc.NC; a.gchunk = sc.gchunk; a.G = pv.G; a.noisepoly = noisepoly; a.sub = sc.sub;
  for (int h = 0; h <= pv.nhyp; h++) a.hyp_begin[h] = pv.hyp_begin[h];
  a.lnO = d_lnO; a.marg = d_marg;
  k_combine<<<(unsigned)((a.BK + 255) / 256), 256, 0, s>>>(a);
  st->launches++;
  return cudaGetLastError();
}

// ======================================================================================
// parameter-estimation grid (bayes.py:912-1055)
// ======================================================================================
__device__ __forceinline__ double model_value(int kind, int reverse, const double* p, double ts, double t0c,
                                              double ts0, const double* cts, int L) {
  // same definitions as k_gen_templates, for one sample; t0c = centre time, ts0 = cts[0]
  double t0 = p[0];
  const bool centre = isinf(t0) && t0 > 0;
  if (centre) t0 = t0c;
  switch (kind) {
    case 0: {
      const double tauexp = p[1], taugauss = p[2], amp = p[3];
      double f = (ts == t0) ? amp : 0.0;
      const bool before = ts < t0, after = ts > t0;
      if (taugauss > 0 && (reverse ? after : before)) {
        const double dt = ts - t0;
        f = amp * exp(-(dt * dt) / (2.0 * (taugauss * taugauss)));
      }
      if (tauexp > 0) {
        if (reverse && before) f = amp * exp((ts - t0) / tauexp);
        if (!reverse && after) f = amp * exp(-(ts - t0) / tauexp);
      }
      return f;
    }
    case 1: {
      const double amp = p[1], tauexp = p[2];
      double f = (ts == t0) ? amp : 0.0;
      if (tauexp > 0) {
        if (reverse && ts < t0) f = amp * exp((ts - t0) / tauexp);
        if (!reverse && ts > t0) f = amp * exp(-(ts - t0) / tauexp);
      }
      return f;
    }
    case 2: {
      if (!centre) t0 = ts0 + p[0];
      int best = 0;
      double bd = fabs(cts[0] - t0);
      for (int i = 1; i < L; i++) {
        const double a = fabs(cts[i] - t0);
        if (a < bd) { bd = a; best = i; }
      }
      return (ts == cts[best]) ? p[1] : 0.0;
    }
    case 3: return (ts > t0) ? p[1] : 0.0;
    case 4: {
      const double sigmag = p[1], tauf = p[2], amp = p[3];
      double f = -1.0 * amp;
      if (ts < t0 - tauf) {
        const double x = ts - (t0 - tauf);
        f = sigmag > 0 ? (-1.0 * amp) * exp(-(x * x) / (2.0 * (sigmag * sigmag))) : 0.0;
      }
      if (ts > t0 + tauf) {
        const double x = ts - (t0 + tauf);
        f = sigmag > 0 ? (-1.0 * amp) * exp(-(x * x) / (2.0 * (sigmag * sigmag))) : 0.0;
      }
      return f;
    }
    case 5: {
      const double sigma = p[1], amp = p[2];
      if (sigma == 0) return ((ts - t0) == (cts[0] - t0)) ? amp : 0.0;
      const double dt = ts - t0;
      return amp * exp(-(dt * dt) / (2.0 * (sigma * sigma)));
    }
  }
  return 0.0;
}

// log_marg_amp_except_final_C (log_marg_amp_full.c:71-134) replayed literally: the Z update
// guard at :107 makes it differ from the exact Gaussian integral for Nm >= 5;
// "same results as the reference" means replaying it.
template <int NM>
__device__ double except_final_logL(const double (&Mu)[NM][NM], const double (&D)[NM], double sigma) {
  constexpr int nm = NM - 1, nmm = NM - 2;
  double sq[NM], c[NM][NM];
#pragma unroll
  for (int i = 0; i < NM; i++)
#pragma unroll
    for (int j = 0; j < NM; j++) c[i][j] = 0.0;
  bool fin = true;
#pragma unroll
  for (int i = 0; i < NM; i++) {
    sq[i] = Mu[i][i];
    c[i][i] = mul_(-2.0, D[i]);
    if (!isfinite(sq[i]) || !isfinite(c[i][i])) fin = false;
#pragma unroll
    for (int j = i + 1; j < NM; j++) {
      c[i][j] = mul_(2.0, Mu[i][j]);
      if (!isfinite(c[i][j])) fin = false;
    }
  }
  if (!fin) return neg_inf();
  double Y = 0.0, Z = 0.0;
#pragma unroll
  for (int i = 0; i < nmm; i++) {
    const double inv = div_(1.0, sq[i]);
    const double inv2 = mul_(0.5, inv), inv4 = mul_(0.25, inv);
#pragma unroll
    for (int j = i; j < NM; j++) {
#pragma unroll
      for (int k = i; k < NM; k++) {
        if ((j != i + 1 && j != nmm) && (k != i + 1 && k != nmm)) Z = sub_(Z, mul_(mul_(c[i][j], c[i][k]), inv4));
        if (j == i) {
          if (k > j && k < nm) sq[k] = sub_(sq[k], mul_(mul_(c[j][k], c[j][k]), inv4));
        } else {
          if (k == i && j < nm) c[j][j] = sub_(c[j][j], mul_(mul_(c[i][j], c[i][k]), inv2));
          else if (k > j) c[j][k] = sub_(c[j][k], mul_(mul_(c[i][j], c[i][k]), inv2));
        }
      }
    }
  }
  const double X = sq[nmm];
#pragma unroll
  for (int i = nmm; i < NM; i++) Y = add_(Y, c[nmm][i]);
  Z = add_(Z, add_(sq[nm], c[nm][nm]));
  double logL = 0.0;
#pragma unroll
  for (int i = 0; i < nm; i++) logL = sub_(logL, mul_(0.5, log(sq[i])));
  const double q = div_(mul_(mul_(0.25, Y), Y), X);
  logL = sub_(logL, div_(mul_(0.5, sub_(Z, q)), mul_(sigma, sigma)));
  logL = add_(logL, mul_((double)nm, log(sigma)));
  logL = add_(logL, mul_(mul_(0.5, (double)nm), add_(BFL_LN2, BFL_LNPI)));
  return logL;
}

struct PeArgs {
  int kind, reverse, L, margbackground;
  const double* cts; const double* clc; const double* params; const double* logprior; const double* polyms;
  const double* subm;   // [NM*NM + NM] background block + data x background, from k_pe_setup
  int64_t G;
  double sigma;
  double* post;
  const double* sgcoef; // Savitzky-Golay taps when the chunk was detrended (bayes.py:1018-1021), else nullptr
  int sgbins;
};
constexpr int PE_MAXL = 200;

// background-only sums shared by every grid point (bayes.py:981-1000), NumPy order
__global__ void k_pe_setup(const double* __restrict__ clc, const double* __restrict__ polyms, int L, int P,
                           double* __restrict__ subm) {
  const int NM = P + 1;
  const int t = threadIdx.x;
  if (t < P * P) {
    const int p = t / P, q = t % P;
    if (q >= p) {
      const double* a = polyms + (size_t)p * L;
      const double* bq = polyms + (size_t)q * L;
      subm[p * NM + q] = pairwise_np(L, [&](int i) { return mul_(a[i], bq[i]); });
    }
  } else if (t < P * P + P) {
    const int p = t - P * P;
    const double* a = polyms + (size_t)p * L;
    subm[NM * NM + p] = pairwise_np(L, [&](int i) { return mul_(a[i], clc[i]); });
  }
}

template <int P>
__global__ void __launch_bounds__(128) k_pe_grid(const PeArgs a) {
  extern __shared__ double smem[];
  constexpr int NM = P + 1;
  const int L = a.L, tid = threadIdx.x;
  double* sts = smem;              // [L]
  double* sd = sts + L;            // [L]
  double* spoly = sd + L;          // [P][L]
  double* sm = spoly + P * L;      // [L][128] template values, transposed
  for (int i = tid; i < L; i += 128) { sts[i] = a.cts[i]; sd[i] = a.clc[i]; }
  if (a.margbackground)
    for (int i = tid; i < P * L; i += 128) spoly[i] = a.polyms[i];
  __syncthreads();
  const int64_t g = (int64_t)blockIdx.x * 128 + tid;
  if (g >= a.G) return;
  double p[4];
  for (int i = 0; i < 4; i++) p[i] = a.params[g * 4 + i];
  const double t0c = sts[L / 2];
  for (int i = 0; i < L; i++) sm[i * 128 + tid] = model_value(a.kind, a.reverse, p, sts[i], t0c, sts[0], sts, L);
  if (a.sgbins > 0) {
    // the light curve was Savitzky-Golay detrended: the model gets the same treatment, m - savitzky_golay(m)
    // (noise.py:176-252: mirrored-difference padding, then the FIR; same arithmetic as k_sg_resid)
    double filt[PE_MAXL];
    const int W = a.sgbins, h = (W - 1) / 2;
    const double y0 = sm[tid], yl = sm[(L - 1) * 128 + tid];
    for (int i = 0; i < L; i++) {
      double fit = 0.0;
      for (int j = 0; j < W; j++) {
        const int t = i + j;
        double v;
        if (t < h) v = y0 - fabs(sm[(h - t) * 128 + tid] - y0);
        else if (t < h + L) v = sm[(t - h) * 128 + tid];
        else v = yl + fabs(sm[(L - 2 - (t - h - L)) * 128 + tid] - yl);
        fit = fma(a.sgcoef[j], v, fit);
      }
      filt[i] = sm[i * 128 + tid] - fit;
    }
    for (int i = 0; i < L; i++) sm[i * 128 + tid] = filt[i];
  }
  double res;
  if (a.margbackground) {
    double Mu[NM][NM], D[NM];
#pragma unroll
    for (int i = 0; i < NM; i++) {
#pragma unroll
      for (int j = 0; j < NM; j++) Mu[i][j] = (j >= i && i < P && j < P) ? a.subm[i * NM + j] : 0.0;
      D[i] = (i < P) ? a.subm[NM * NM + i] : 0.0;
    }
#pragma unroll
    for (int j = 0; j < P; j++) {
      const double* pj = spoly + j * L;
      Mu[j][P] = pairwise_np(L, [&](int i) { return mul_(sm[i * 128 + tid], pj[i]); });   // bayes.py:1027-1028
    }
    Mu[P][P] = pairwise_np(L, [&](int i) { const double m = sm[i * 128 + tid]; return mul_(m, m); });
    D[P] = pairwise_np(L, [&](int i) { return mul_(sd[i], sm[i * 128 + tid]); });
    res = except_final_logL<NM>(Mu, D, a.sigma);
  } else {
    // log_likelihood_ratio, general.pyx:659-691: -sum(m^2 - 2 d m) / (2 sigma^2)
    const double s = pairwise_np(L, [&](int i) {
      const double m = sm[i * 128 + tid];
      return sub_(mul_(m, m), mul_(mul_(2.0, sd[i]), m));
    });
    res = div_(-s, mul_(2.0, mul_(a.sigma, a.sigma)));
  }
  a.post[g] = res + a.logprior[g];
}

cudaError_t launch_pe_grid(int kind, int reverse, const double* d_cts, const double* d_clc, int L,
                           const double* d_params, const double* d_logprior, int64_t G, int P,
                           const double* d_polyms, double sigma, int margbackground, double* d_post,
                           const double* d_sgcoef, int sgbins, cudaStream_t s, LaunchStats* st) {
  PeArgs a;
  a.sgcoef = d_sgcoef; a.sgbins = d_sgcoef ? sgbins : 0;
  a.kind = kind; a.reverse = reverse; a.L = L; a.margbackground = margbackground;
  a.cts = d_cts; a.clc = d_clc; a.params = d_params; a.logprior = d_logprior; a.polyms = d_polyms;
  a.G = G; a.sigma = sigma; a.post = d_post;
  double* d_subm = nullptr;
  cudaError_t e = cudaSuccess;
  if (margbackground) {
    e = cudaMallocAsync((void**)&d_subm, sizeof(double) * (size_t)((P + 1) * (P + 1) + (P + 1)), s);
    if (e != cudaSuccess) return e;
    k_pe_setup<<<1, 64, 0, s>>>(d_clc, d_polyms, L, P, d_subm);
    st->launches++;
  }
  a.subm = d_subm;
  const size_t smem = sizeof(double) * ((size_t)L * 2 + (size_t)P * L + (size_t)L * 128);
  const unsigned grid = (unsigned)((G + 127) / 128);
#define BFL_PE(P_)                                                                                      \
  case P_:                                                                                              \
    e = cudaFuncSetAttribute(k_pe_grid<P_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
    if (e == cudaSuccess) k_pe_grid<P_><<<grid, 128, smem, s>>>(a);                                     \
    break;
  switch (P) {
    BFL_PE(1) BFL_PE(2) BFL_PE(3) BFL_PE(4) BFL_PE(5) BFL_PE(6)
    default: e = cudaErrorInvalidValue;
  }
#undef BFL_PE
  st->launches++;
  if (d_subm) cudaFreeAsync(d_subm, s);
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

// ======================================================================================
// logtrapz (general.pyx:576-607) along the leading axis of lx[n][m]
// ======================================================================================
__global__ void k_logtrapz(const double* __restrict__ lx, const double* __restrict__ t, int64_t n, int64_t m,
                           double* __restrict__ out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  double B = neg_inf();
  for (int64_t i = 0; i + 1 < n; i++) {
    double z = logplus_dev(lx[i * m + j], lx[(i + 1) * m + j]);
    z = z + log(t[i + 1] - t[i]) - BFL_LN2;
    B = logplus_dev(z, B);
  }
  out[j] = B;
}

cudaError_t launch_logtrapz(const double* d_lx, const double* d_t, int64_t n, int64_t m, double* d_out,
                            cudaStream_t s, LaunchStats* st) {
  k_logtrapz<<<(unsigned)((m + 127) / 128), 128, 0, s>>>(d_lx, d_t, n, m, d_out);
  st->launches++;
  return cudaGetLastError();
}

// ======================================================================================
// thresholding (find.py:15-52, 915-1019)
// ======================================================================================
__global__ void k_threshold_edges(const double* __restrict__ lnO, int64_t n, double thresh, int64_t* starts,
                                  int64_t* stops, int* counts, int cap) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool cur = lnO[i] > thresh;
  if (!cur) return;
  const bool prev = (i > 0) && (lnO[i - 1] > thresh);
  const bool next = (i + 1 < n) && (lnO[i + 1] > thresh);
  if (!prev) { const int s = atomicAdd(&counts[0], 1); if (s < cap) starts[s] = i; }
  if (!next) { const int s = atomicAdd(&counts[1], 1); if (s < cap) stops[s] = i + 1; }
}

cudaError_t launch_threshold_edges(const double* d_lnO, int64_t n, double thresh, int64_t* d_starts,
                                   int64_t* d_stops, int* d_counts, int cap, cudaStream_t s, LaunchStats* st) {
  if (n <= 0) return cudaSuccess;
  k_threshold_edges<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_lnO, n, thresh, d_starts, d_stops, d_counts, cap);
  st->launches++;
  return cudaGetLastError();
}

// number of region starts per curve (a start = above threshold and predecessor not above)
// one CTA per curve: contiguous_regions (find.py:15-52) counts the samples that OPEN a region
__global__ void __launch_bounds__(256) k_count_regions(const double* __restrict__ lnO, int B, int64_t n, double thresh,
                                                       unsigned long long* counts, unsigned long long* totals) {
  const int b = blockIdx.x;
  const double* row = lnO + (size_t)b * n;
  int cnt = 0;
  for (int64_t k = threadIdx.x; k < n; k += 256) {
    const bool up = row[k] > thresh;
    const bool prev = (k > 0) && (row[k - 1] > thresh);
    cnt += (up && !prev) ? 1 : 0;
  }
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
  __shared__ int part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < 8; w++) tot += part[w];
    counts[b] = (unsigned long long)tot;
    if (totals && tot > 0) {                      // [0] regions in the batch, [1] curves with at least one
      atomicAdd(&totals[0], (unsigned long long)tot);
      atomicAdd(&totals[1], 1ULL);
    }
  }
}

cudaError_t launch_count_regions(const double* d_lnO, int B, int64_t n, double thresh, int64_t* d_counts,
                                 int64_t* d_totals, cudaStream_t s, LaunchStats* st) {
  k_count_regions<<<(unsigned)B, 256, 0, s>>>(d_lnO, B, n, thresh, (unsigned long long*)d_counts,
                                              (unsigned long long*)d_totals);
  st->launches++;
  return cudaGetLastError();
}

// Regions of a whole batch in one launch (contiguous_regions, find.py:15-52, for every curve): the thread of a
// sample that OPENS a region (above threshold, predecessor not) walks to its end -- regions are a few samples
// long and rare -- and emits one record {curve, start, stop, argmax, max} (first maximum, as np.argmax);
// per-curve counts and batch totals as k_count_regions.  Record order is arbitrary (atomic slot counter).
__global__ void k_batch_regions(const double* __restrict__ lnO, int B, int64_t n, double thresh, RegionRec* recs, int cap,
                                unsigned long long* counts, unsigned long long* totals) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * B) return;
  const int64_t b = i / n, k = i - b * n;
  if (!(lnO[i] > thresh)) return;
  if (k > 0 && lnO[i - 1] > thresh) return;
  const double* row = lnO + b * n;
  double best = row[k];
  int64_t bi = k, e = k + 1;
  while (e < n && row[e] > thresh) {
    if (row[e] > best) { best = row[e]; bi = e; }
    e++;
  }
  const unsigned long long before = atomicAdd(&counts[b], 1ULL);
  atomicAdd(&totals[0], 1ULL);
  if (before == 0ULL) atomicAdd(&totals[1], 1ULL);
  const unsigned long long slot = atomicAdd(&totals[2], 1ULL);
  if (slot < (unsigned long long)cap) {
    RegionRec r;
    r.curve = (int)b; r.start = (int)k; r.stop = (int)e; r.argmax = (int)bi; r.maxval = best;
    recs[slot] = r;
  }
}

cudaError_t launch_batch_regions(const double* d_lnO, int B, int64_t n, double thresh, RegionRec* d_recs, int cap,
                                 int64_t* d_counts, int64_t* d_totals, cudaStream_t s, LaunchStats* st) {
  const int64_t tot = n * B;
  k_batch_regions<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(d_lnO, B, n, thresh, d_recs, cap,
                                                                (unsigned long long*)d_counts, (unsigned long long*)d_totals);
  st->launches++;
  return cudaGetLastError();
}

// one warp per region: maximum and its FIRST index (np.argmax, find.py:1012-1015).
// NaN entries are skipped (a region above threshold always holds a non-NaN value).
__global__ void k_region_max(const double* __restrict__ lnO, const int64_t* __restrict__ segs, int nsegs,
                             double* maxval, int64_t* maxidx) {
  const int r = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  if (r >= nsegs) return;
  const int lane = threadIdx.x & 31;
  const int64_t s0 = segs[2 * r], s1 = segs[2 * r + 1];
  double best = neg_inf();
  int64_t bi = INT64_MAX;
  for (int64_t i = s0 + lane; i < s1; i += 32) {
    const double v = lnO[i];
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_down_sync(0xffffffffu, best, o);
    const int64_t oi = __shfl_down_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (lane == 0) { maxval[r] = best; maxidx[r] = (bi == INT64_MAX) ? s0 : bi; }
}

cudaError_t launch_region_max(const double* d_lnO, const int64_t* d_segs, int nsegs, double* d_maxval,
                              int64_t* d_maxidx, cudaStream_t s, LaunchStats* st) {
  if (nsegs <= 0) return cudaSuccess;
  k_region_max<<<(nsegs + 3) / 4, 128, 0, s>>>(d_lnO, d_segs, nsegs, d_maxval, d_maxidx);
  st->launches++;
  return cudaGetLastError();
}


// ======================================================================================
// noise-sigma estimation on device: noise/noise.py:10-103, 176-252,
// data/data.py:184-203, 377-409, stats/bayes.py:235-246
// ======================================================================================
// Savitzky-Golay detrend: resid = y - fit, fit[i] = sum_j m[j] * ypad[i + j], with the reference's
// mirrored-difference edge padding (noise.py:247-251).  m = pinv(b)[0] comes from the caller.
// One CTA per 256 consecutive samples of one curve: the padded window (mirrored differences at both ends,
// noise.py:236-240) is staged in shared memory once, every thread then takes its W taps from there in the
// reference's order (np.convolve(m[::-1], y, 'valid')), so the result is bit-identical to a thread-per-sample
// loop over global memory -- which took 29 us for 128 curves, a seventh of a whole detector step.
constexpr int SG_TPB = 128, SG_KT = 4;       // 128 threads x 4 consecutive samples each = 512 samples per CTA
__global__ void __launch_bounds__(SG_TPB) k_sg_resid(const double* __restrict__ y, int64_t N, int B,
                                                     const double* __restrict__ m, int W, double* __restrict__ resid) {
  extern __shared__ double smem[];
  constexpr int TS = SG_TPB * SG_KT;
  double* sv = smem;                   // [TS + W - 1] padded series, positions i0 .. i0 + TS + W - 2
  double* sm = smem + TS + W - 1;      // [W] filter taps
  const int b = blockIdx.y;
  const int64_t i0 = (int64_t)blockIdx.x * TS;
  const double* yy = y + (size_t)b * N;
  const int h = (W - 1) / 2;
  const double y0 = yy[0], yl = yy[N - 1];
  for (int e = threadIdx.x; e < TS + W - 1; e += SG_TPB) {
    const int64_t t = i0 + e;          // index into the padded series
    double v = 0.0;
    if (t < h) v = y0 - fabs(yy[h - t] - y0);
    else if (t < h + N) v = yy[t - h];
    else if (t < N + 2 * h) v = yl + fabs(yy[N - 2 - (t - h - N)] - yl);
    sv[e] = v;
  }
  for (int j = threadIdx.x; j < W; j += SG_TPB) sm[j] = m[j];
  __syncthreads();
  // SG_KT consecutive samples per thread: every padded value is read once per thread instead of once per sample,
  // every tap once per SG_KT samples; each sample's sum still runs over the taps in the reference's order
  const int base = threadIdx.x * SG_KT;
  double fit[SG_KT];
#pragma unroll
  for (int q = 0; q < SG_KT; q++) fit[q] = 0.0;
  double win[SG_KT];
#pragma unroll
  for (int q = 0; q < SG_KT - 1; q++) win[q + 1] = sv[base + q];
  for (int j = 0; j < W; j++) {
#pragma unroll
    for (int q = 0; q < SG_KT - 1; q++) win[q] = win[q + 1];
    win[SG_KT - 1] = sv[base + j + SG_KT - 1];
    const double c = sm[j];
#pragma unroll
    for (int q = 0; q < SG_KT; q++) fit[q] = fma(c, win[q], fit[q]);
  }
#pragma unroll
  for (int q = 0; q < SG_KT; q++) {
    const int64_t i = i0 + base + q;
    if (i < N) resid[(size_t)b * N + i] = sv[base + q + h] - fit[q];
  }
}

// 'powerspectrum' estimator: sigma^2 = mean of the upper `estfrac` of the one-sided periodogram
// times fs/2 (noise.py:36-40) = sum_k c_k |X_k|^2 / (2 n cnt), c_k = 2 except DC and Nyquist.
__global__ void __launch_bounds__(256) k_ps_sigma(const double2* __restrict__ X, int64_t N, int64_t L, int64_t k0,
                                                  double* __restrict__ sk) {
  const int b = blockIdx.x;
  const double2* x = X + (size_t)b * L;
  double s = 0.0;
  for (int64_t k = k0 + threadIdx.x; k < L; k += 256) {
    const double2 v = x[k];
    const double p = fma(v.x, v.x, v.y * v.y);
    const bool single = (k == 0) || ((N % 2 == 0) && (k == N / 2));
    s += single ? p : 2.0 * p;
  }
  __shared__ double red[256];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) sk[b] = sqrt(red[0] / (2.0 * (double)N * (double)(L - k0)));
}

// 'tailveto' estimator, part 1: N-bin histogram of one curve with NumPy's bin assignment
// (numpy.histogram uniform-bin path incl. its one-ulp edge corrections), one block per curve.
__device__ __forceinline__ double tv_edge(int64_t i, int64_t N, double lo, double hi, double step) {
  return (i == N) ? hi : __dadd_rn(__dmul_rn((double)i, step), lo);    // np.linspace: arange*step + start, last = stop
}

__global__ void __launch_bounds__(256) k_tv_hist(const double* __restrict__ d, int64_t N, int* __restrict__ counts,
                                                 double* __restrict__ range) {
  const int b = blockIdx.x;
  const double* x = d + (size_t)b * N;
  __shared__ double smin[256], smax[256];
  double mn = INFINITY, mx = -INFINITY;
  bool bad = false;
  for (int64_t i = threadIdx.x; i < N; i += 256) {
    const double v = x[i];
    if (!isfinite(v)) bad = true;
    mn = fmin(mn, v); mx = fmax(mx, v);
  }
  smin[threadIdx.x] = bad ? nan("") : mn;
  smax[threadIdx.x] = mx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const double a = smin[threadIdx.x], c = smin[threadIdx.x + o];
      smin[threadIdx.x] = (isnan(a) || isnan(c)) ? nan("") : fmin(a, c);
      smax[threadIdx.x] = fmax(smax[threadIdx.x], smax[threadIdx.x + o]);
    }
    __syncthreads();
  }
  double lo = smin[0], hi = smax[0];
  if (isnan(lo)) {   // NaN / inf in the data: np.histogram raises; report NaN
    if (threadIdx.x == 0) { range[2 * b] = nan(""); range[2 * b + 1] = nan(""); }
    return;
  }
  if (lo == hi) { lo -= 0.5; hi += 0.5; }      // numpy _get_outer_edges
  if (threadIdx.x == 0) { range[2 * b] = lo; range[2 * b + 1] = hi; }
  const double step = (hi - lo) / (double)N;
  const double denom = hi - lo;
  int* c = counts + (size_t)b * N;
  for (int64_t i = threadIdx.x; i < N; i += 256) {
    const double v = x[i];
    int64_t idx = (int64_t)(__dmul_rn(__ddiv_rn(v - lo, denom), (double)N));
    if (idx == N) idx -= 1;
    if (v < tv_edge(idx, N, lo, hi, step)) idx -= 1;
    if (v >= tv_edge(idx + 1, N, lo, hi, step) && idx != N - 1) idx += 1;
    atomicAdd(c + idx, 1);
  }
}

// part 2 (one WARP per curve): density, cumulative distribution, np.unique + scipy interp1d at
// 0.5 -/+ cp/2 (noise.py:83-101).  The cumulative sum must run in np.cumsum's order (the unique test compares
// consecutive sums for equality), so it stays sequential -- but everything that does not depend on the running
// sum (bin edges, the two divisions of the density, the bin centre) is computed by the 32 lanes for 32 bins at
// a time with coalesced loads, and the serial part is one shuffle + one add per bin, the same in every lane.
__global__ void __launch_bounds__(128) k_tv_finish(const int* __restrict__ counts, const double* __restrict__ range,
                                                   int64_t N, int B, double sigma, double cp, double* __restrict__ sk) {
  const int b = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const double lo = range[2 * b], hi = range[2 * b + 1];
  if (isnan(lo)) { if (lane == 0) sk[b] = nan(""); return; }
  const int* c = counts + (size_t)b * N;
  const double step = (hi - lo) / (double)N;
  const double w0 = tv_edge(1, N, lo, hi, step) - tv_edge(0, N, lo, hi, step);
  const double tgt[2] = {0.5 - cp / 2.0, 0.5 + cp / 2.0};
  double val[2] = {nan(""), nan("")};
  int state[2] = {0, 0};            // 0 searching, 1 waiting for the second unique point, 2 done, 3 out of range
  double cs = 0.0, xl = 0.0, yl = 0.0, xp = 0.0, yp = 0.0;
  int nu = 0;
  for (int64_t base = 0; base < N; base += 32) {
    const int64_t i = base + lane;
    double term = 0.0, ymid = 0.0;
    if (i < N) {
      const double e0 = tv_edge(i, N, lo, hi, step), e1 = tv_edge(i + 1, N, lo, hi, step);
      const double ni = __ddiv_rn(__ddiv_rn((double)c[i], e1 - e0), (double)N);   // n / db / n.sum()
      term = __dmul_rn(ni, w0);
      ymid = (e0 + e1) / 2.0;
    }
    const int cnt = (int)min((int64_t)32, N - base);
    for (int t = 0; t < cnt; t++) {
      cs = __dadd_rn(cs, __shfl_sync(0xffffffffu, term, t));
      if (nu == 0 || cs != xl) {       // a new unique value of the cumulative distribution (first index kept)
        xp = xl; yp = yl;
        xl = cs; yl = __shfl_sync(0xffffffffu, ymid, t);
        nu++;
#pragma unroll
        for (int q = 0; q < 2; q++) {
          if (state[q] == 1) {           // target sits at/below the first point: interpolate on points 0 and 1
            val[q] = (yl - yp) / (xl - xp) * (tgt[q] - xp) + yp;
            state[q] = 2;
          } else if (state[q] == 0 && xl >= tgt[q]) {
            if (nu == 1) state[q] = (tgt[q] < xl) ? 3 : 1;     // below the interpolation range -> error
            else { val[q] = (yl - yp) / (xl - xp) * (tgt[q] - xp) + yp; state[q] = 2; }
          }
        }
      }
    }
    if (state[0] >= 2 && state[1] >= 2) break;      // both quantiles settled (warp-uniform)
  }
  if (lane == 0) sk[b] = (state[0] == 2 && state[1] == 2) ? (val[1] - val[0]) / (2.0 * sigma) : nan("");
}

cudaError_t launch_sg_resid(const double* d_y, int64_t N, int B, const double* d_m, int W, double* d_resid,
                            cudaStream_t s, LaunchStats* st) {
  const int ts = SG_TPB * SG_KT;
  const dim3 grid((unsigned)((N + ts - 1) / ts), (unsigned)B);
  const size_t smem = sizeof(double) * (size_t)(ts + 2 * W - 1);
  k_sg_resid<<<grid, SG_TPB, smem, s>>>(d_y, N, B, d_m, W, d_resid);
  st->launches++;
  return cudaGetLastError();
}

cudaError_t launch_ps_sigma(const double2* d_X, int64_t N, int B, double estfrac, double* d_sk, cudaStream_t s,
                            LaunchStats* st) {
  const int64_t L = N / 2 + 1;
  const int64_t k0 = (int64_t)floor((1.0 - estfrac) * (double)L);
  k_ps_sigma<<<B, 256, 0, s>>>(d_X, N, L, k0, d_sk);
  st->launches++;
  return cudaGetLastError();
}

cudaError_t launch_tv_sigma(const double* d_resid, int64_t N, int B, double sigma, int* d_counts, double* d_range,
                            double* d_sk, cudaStream_t s, LaunchStats* st) {
  cudaError_t e = cudaMemsetAsync(d_counts, 0, sizeof(int) * (size_t)N * B, s);
  if (e != cudaSuccess) return e;
  k_tv_hist<<<B, 256, 0, s>>>(d_resid, N, d_counts, d_range);
  st->launches++;
  k_tv_finish<<<(B + 3) / 4, 128, 0, s>>>(d_counts, d_range, N, B, sigma, erf(sigma / sqrt(2.0)), d_sk);
  st->launches++;
  return cudaGetLastError();
}

// ======================================================================================
// FP64 FMA throughput probe: 16 independent chains per thread
// ======================================================================================
__global__ void __launch_bounds__(256) k_fp64_probe(double* sink, int iters) {
  double a[16];
  const double x = 1.0 + 1e-9 * threadIdx.x, y = 1e-12 * (blockIdx.x + 1);
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = (double)i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int i = 0; i < 16; i++) a[i] = fma(a[i], x, y);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += a[i];
  if (s == 123.456) sink[0] = s;   // never true; keeps the chains alive
}

cudaError_t launch_fp64_probe(double* d_sink, int iters, int blocks, cudaStream_t s, LaunchStats* st) {
  k_fp64_probe<<<blocks, 256, 0, s>>>(d_sink, iters);
  st->launches++;
  return cudaGetLastError();
}

}  // namespace bfl

// ======================================================================================
// Injection sweep on the device (scripts/flare_detection_efficiency.py:399-651)
// ======================================================================================
namespace bfl {

// Philox4x32-10 (Salmon et al. 2011): counter-based, so realisation r gets the same stream whatever
// the batch or rank that simulates it.
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
  const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const unsigned hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const unsigned hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ double u53(unsigned a, unsigned b) {     // uniform in [0, 1)
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

struct SimArgs {
  unsigned long long seed;
  long long first;          // index of the batch's first realisation
  int B, N, H;
  double dt, nstd, sinusoid_amp, snr_lo, snr_hi;
  int inject;
  double* clc;              // [B][N]
  double* truth;            // [B][4]  idx, snr, taugauss, tauexp
};

// One CTA per light curve: white noise (Box-Muller on Philox), optional low-frequency sinusoid, one
// injected flare with tau_gauss <= tau_exp scaled to a uniform signal-to-noise ratio
// (flare_detection_efficiency.py:262-289, 321-335, 416-434, 483-487), median removed as
// Lightcurve.dcoffset does (data.py:271-277; block-wide bitonic sort).
__global__ void __launch_bounds__(256) k_sweep_simulate(const SimArgs a) {
  extern __shared__ double smem[];
  const int N = a.N, tid = threadIdx.x, T = blockDim.x;
  int NP = 1;
  while (NP < N) NP <<= 1;
  double* x = smem;           // [N]
  double* srt = smem + N;     // [NP]
  __shared__ double red[256];
  __shared__ double par[8];
  const long long real = a.first + blockIdx.x;
  const uint2 key = make_uint2((unsigned)a.seed, (unsigned)(a.seed >> 32));
  const unsigned r_lo = (unsigned)real, r_hi = (unsigned)((unsigned long long)real >> 32);

  for (int p = tid; 2 * p < N; p += T) {
    const uint4 v = philox4x32(make_uint4((unsigned)p, r_lo, r_hi, 0u), key);
    const double u1 = u53(v.x, v.y) + 5.5511151231257827e-17, u2 = u53(v.z, v.w);
    const double r = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    x[2 * p] = a.nstd * r * c;
    if (2 * p + 1 < N) x[2 * p + 1] = a.nstd * r * s;
  }
  if (tid == 0) {
    // per-realisation parameters from a separate counter plane (w = 1)
    unsigned q = 0;
    auto next = [&](double& u_a, double& u_b) {
      const uint4 v = philox4x32(make_uint4(q++, r_lo, r_hi, 1u), key);
      u_a = u53(v.x, v.y); u_b = u53(v.z, v.w);
    };
    double ua, ub;
    next(ua, ub);
    const double amp = ua * a.sinusoid_amp * a.nstd;
    const double f = 1.0 / (31.0 * 86400.0) + ub * (1.0 / (2.0 * 86400.0) - 1.0 / (31.0 * 86400.0));
    next(ua, ub);
    const double phase = 2.0 * ua;                         // in units of pi
    int idx = a.H + (int)(ub * (double)(N - 2 * a.H - 1));  // randint(h, N-h-1)
    idx = min(idx, N - a.H - 2);
    next(ua, ub);
    double te = 1800.0 + ua * 9000.0;
    double tg = ub * 5400.0;
    if (a.inject != 3) {
      while (tg > te) { next(ua, ub); tg = ua * 5400.0; }   // redrawn until tau_gauss <= tau_exp (script :321-335)
    } else {
      // transit (script :291-312): tauf U[1800, 14400], sigmag U[0, 10800], both redrawn until tauf + 4 sigmag <= 12 h;
      // recorded as (p1, p2) = (sigmag, tauf)
      te = 1800.0 + ua * 12600.0;
      tg = ub * 10800.0;
      while (te + 4.0 * tg > 43200.0) { next(ua, ub); te = 1800.0 + ua * 12600.0; tg = ub * 10800.0; }
    }
    next(ua, ub);
    const double snr = a.snr_lo + ua * (a.snr_hi - a.snr_lo);
    par[0] = amp; par[1] = f; par[2] = phase; par[3] = (double)idx; par[4] = te; par[5] = tg; par[6] = snr;
  }
  __syncthreads();
  const double amp = par[0], f = par[1], phase = par[2], te = par[4], tg = par[5], snr = par[6];
  const int idx = (int)par[3];
  if (a.sinusoid_amp > 0.0)
    for (int i = tid; i < N; i += T) x[i] += amp * sinpi(2.0 * f * ((double)i * a.dt) + phase);
  if (a.inject == 1) {
    const int lo = max(0, idx - 4 * a.H), hi = min(N, idx + 12 * a.H);   // the template is negligible outside
    auto flare = [&](int i) {                                           // Flare.model, models/model.py:197-252
      const double t = (double)(i - idx) * a.dt;
      if (i == idx) return 1.0;
      if (i < idx) return tg > 0.0 ? exp(-t * t / (2.0 * tg * tg)) : 0.0;
      return te > 0.0 ? exp(-t / te) : 0.0;
    };
    double s2 = 0.0;
    for (int i = lo + tid; i < hi; i += T) { const double m = flare(i); s2 = fma(m, m, s2); }
    red[tid] = s2;
    __syncthreads();
    for (int o = T / 2; o > 0; o >>= 1) { if (tid < o) red[tid] += red[tid + o]; __syncthreads(); }
    const double scale = snr * a.nstd / sqrt(red[0]);
    for (int i = lo + tid; i < hi; i += T) x[i] += flare(i) * scale;
  }
  __syncthreads();
  if (a.inject >= 2) {
    // base curve of the script's recipe (noise + sinusoid; the injection is scaled by the noise estimate of THIS
    // curve and added by k_sweep_inject): no median removal, the script does none for simulated data
    double* out = a.clc + (size_t)blockIdx.x * N;
    for (int i = tid; i < N; i += T) out[i] = x[i];
    if (tid == 0) {
      double* t = a.truth + (size_t)blockIdx.x * 4;
      t[0] = (double)idx; t[1] = snr; t[2] = tg; t[3] = te;
    }
    return;
  }
  // median: bitonic sort of a padded copy
  for (int i = tid; i < NP; i += T) srt[i] = i < N ? x[i] : __longlong_as_double(0x7ff0000000000000LL);
  __syncthreads();
  for (int k = 2; k <= NP; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < NP; i += T) {
        const int l = i ^ j;
        if (l > i) {
          const double vi = srt[i], vl = srt[l];
          const bool up = (i & k) == 0;
          if ((vi > vl) == up) { srt[i] = vl; srt[l] = vi; }
        }
      }
      __syncthreads();
    }
  const double med = (N & 1) ? srt[N / 2] : 0.5 * (srt[N / 2 - 1] + srt[N / 2]);
  double* out = a.clc + (size_t)blockIdx.x * N;
  for (int i = tid; i < N; i += T) out[i] = x[i] - med;
  if (tid == 0) {
    double* t = a.truth + (size_t)blockIdx.x * 4;
    t[0] = a.inject ? (double)idx : -1.0; t[1] = snr; t[2] = tg; t[3] = te;
  }
}

cudaError_t launch_sweep_simulate(unsigned long long seed, long long first, int B, int N, int H, double dt, double nstd,
                                  double sinusoid_amp, double snr_lo, double snr_hi, int inject, double* d_clc,
                                  double* d_truth, cudaStream_t s, LaunchStats* st) {
  SimArgs a{seed, first, B, N, H, dt, nstd, sinusoid_amp, snr_lo, snr_hi, inject, d_clc, d_truth};
  int NP = 1;
  while (NP < N) NP <<= 1;
  const size_t smem = sizeof(double) * (size_t)(N + NP);
  cudaError_t e = cudaFuncSetAttribute(k_sweep_simulate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_sweep_simulate<<<B, 256, smem, s>>>(a);
  st->launches++;
  return cudaGetLastError();
}

// The script's injection (scripts/flare_detection_efficiency.py:444-487): the model at amplitude 1 on the curve's
// time stamps, rescaled to the drawn signal-to-noise ratio with the noise sigma ESTIMATED from the base curve
// (sk, tail veto of the Savitzky-Golay residual): clc += inj * snr / ((1/sk) sqrt(sum inj^2)).  One CTA per curve.
// kind 1: Flare (p1, p2) = (tau_gauss, tau_exp), models/model.py:197-252; kind 2: Transit (p1, p2) = (sigma_g, tau_f),
// models/model.py:387-436 (amp = 1: -1 inside |t - t0| <= tau_f, Gaussian wings outside).
__global__ void __launch_bounds__(256) k_sweep_inject(int kind, int N, double dt, const double* __restrict__ sk,
                                                      const double* truth, double* __restrict__ clc, int clear_idx,
                                                      double* truth_rw) {
  const int b = blockIdx.x, tid = threadIdx.x;
  __shared__ double red[256];
  const double* t = truth + (size_t)b * 4;
  const int idx = (int)t[0];
  if (idx < 0) return;
  const double snr = t[1], p1 = t[2], p2 = t[3];
  const double t0 = (double)idx * dt;
  auto model = [&](int i) {
    const double ts = (double)i * dt;
    if (kind == 1) {
      if (i == idx) return 1.0;
      const double d = ts - t0;
      if (i < idx) return p1 > 0.0 ? exp(-(d * d) / (2.0 * (p1 * p1))) : 0.0;
      return p2 > 0.0 ? exp(-d / p2) : 0.0;
    }
    double f = -1.0;
    if (ts < t0 - p2) { const double x = ts - (t0 - p2); f = p1 > 0.0 ? -exp(-(x * x) / (2.0 * (p1 * p1))) : 0.0; }
    if (ts > t0 + p2) { const double x = ts - (t0 + p2); f = p1 > 0.0 ? -exp(-(x * x) / (2.0 * (p1 * p1))) : 0.0; }
    return f;
  };
  double s2 = 0.0;
  for (int i = tid; i < N; i += 256) { const double m = model(i); s2 = fma(m, m, s2); }
  red[tid] = s2;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (tid < o) red[tid] += red[tid + o]; __syncthreads(); }
  const double injsnr = (1.0 / sk[b]) * sqrt(red[0]);
  const double scale = snr / injsnr;
  double* x = clc + (size_t)b * N;
  for (int i = tid; i < N; i += 256) x[i] += model(i) * scale;
  // script :616-623: with a transit injected, EVERY above-threshold segment of the flare detector is a false positive
  // ("a transit or impulse was detected, so anything is a false positive"): the scoring kernel gets no injection index
  if (kind == 2 && clear_idx) {
    __syncthreads();
    if (tid == 0) truth_rw[(size_t)b * 4] = -1.0;
  }
}

cudaError_t launch_sweep_inject(int kind, int B, int N, double dt, const double* d_sk, double* d_truth, double* d_clc,
                                cudaStream_t s, LaunchStats* st) {
  k_sweep_inject<<<B, 256, 0, s>>>(kind, N, dt, d_sk, d_truth, d_clc, 1, d_truth);
  st->launches++;
  return cudaGetLastError();
}

// One warp per curve: above-threshold runs expanded by +-2 samples and merged when they touch
// (flare_detection_efficiency.py:521-580); the region holding the injection index is a detection, every
// other region a false positive (:586-612).  Histograms over signal-to-noise ratio as np.histogram.
// The warp reads 32 samples at a time (coalesced) and ballots them into a bit mask; the run/region state
// machine works on the masks with bit scans (uniform in every lane), so noise-only stretches cost one
// load and one ballot per 32 samples.
__global__ void __launch_bounds__(128) k_sweep_score(const double* __restrict__ lnO, int B, int64_t nk, int h,
                                                     const double* __restrict__ truth, double thresh, int nbins,
                                                     double snr_max, unsigned long long* injected,
                                                     unsigned long long* detected, unsigned long long* counters) {
  const int b = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const double* row = lnO + (size_t)b * nk;
  const int64_t inj = (int64_t)truth[4 * b] - h;           // injection index in the edge-trimmed series
  const bool has_inj = truth[4 * b] >= 0.0;
  int64_t rs = -1, re = -1, run_start = 0;
  int nfalse = 0, nreg = 0;
  bool hit = false, in_run = false;
  auto close_region = [&]() {
    if (rs < 0) return;
    nreg++;
    if (has_inj && inj >= rs && inj < re) hit = true; else nfalse++;
  };
  auto emit_run = [&](int64_t a, int64_t e) {              // above-threshold run [a, e)
    const int64_t xs = max((int64_t)0, a - 2), xe = min(nk, e + 2);
    if (rs >= 0 && xs <= re) re = max(re, xe);
    else { close_region(); rs = xs; re = xe; }
  };
  for (int64_t kb = 0; kb < nk; kb += 32) {
    const int64_t k = kb + lane;
    const unsigned m = __ballot_sync(0xffffffffu, k < nk && row[k] > thresh);
    if (m == 0u && !in_run) continue;
    int pos = 0;
    while (pos < 32) {
      const unsigned from = (pos == 0) ? 0xffffffffu : (0xffffffffu << pos);
      if (in_run) {
        const unsigned z = ~m & from;
        if (z == 0u) break;                                // the run continues into the next chunk
        const int q = __ffs(z) - 1;
        emit_run(run_start, kb + q);
        in_run = false;
        pos = q + 1;
      } else {
        const unsigned o = m & from;
        if (o == 0u) break;
        const int q = __ffs(o) - 1;
        run_start = kb + q;
        in_run = true;
        pos = q + 1;
      }
    }
  }
  if (in_run) emit_run(run_start, nk);
  close_region();
  if (lane != 0) return;
  if (has_inj) {
    const double snr = truth[4 * b + 1];
    int bin = (int)floor(snr / snr_max * nbins);
    if (snr == snr_max) bin = nbins - 1;                   // np.histogram: the last edge is inclusive
    if (bin >= 0 && bin < nbins) {
      atomicAdd(&injected[bin], 1ULL);
      if (hit) atomicAdd(&detected[bin], 1ULL);
    }
  }
  if (nfalse) atomicAdd(&counters[0], (unsigned long long)nfalse);
  if (nreg) atomicAdd(&counters[1], (unsigned long long)nreg);
}

cudaError_t launch_sweep_score(const double* d_lnO, int B, int64_t nk, int h, const double* d_truth, double thresh,
                               int nbins, double snr_max, int64_t* d_injected, int64_t* d_detected, int64_t* d_counters,
                               cudaStream_t s, LaunchStats* st) {
  k_sweep_score<<<(B + 3) / 4, 128, 0, s>>>(d_lnO, B, nk, h, d_truth, thresh, nbins, snr_max,
                                                (unsigned long long*)d_injected, (unsigned long long*)d_detected,
                                                (unsigned long long*)d_counters);
  st->launches++;
  return cudaGetLastError();
}

}  // namespace bfl
