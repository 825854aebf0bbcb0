// bfl_wsplit.cu -- weighted path (per-sample variances), SPLIT form for sm_100a: two kernels per chunk of curves.
//
// Mathematics as in k_window_wsep (bfl_kernels.cu; reference: bayesflare/stats/bayes.py:312-407 cross terms with
// v_j = sk^2 + cle_j^2, general.pyx:143-312, log_marg_amp_full.c:3-68): per sample the P x P system
// A = Q U Q^T = L L^T of the orthonormal background basis and yc = L^-1 Q U d; per SHAPE m (rise or decay part of a
// flare template, or a dense noise-model template)  f = Q U m, e = m U m, r0 = m U d  ->  y = L^-1 f,
// s' = e - |y|^2, r' = r0 - y.yc; per template s = s'_A + s'_B - 2 yA.yB, r = r'_A + r'_B, z = r / sqrt(s) and the
// grid's log-sum-exp with weight 1/sqrt(s).
//
// Every window sum above is a CORRELATION of one of two streams, u = 1/v or u d, with a vector that is fixed per
// plan (q_p q_q, m q_p, m^2 on u; q_p, m on u d -- built by bfl_plan_create, PlanView::wc_*).  So the path splits
// exactly like the constant-variance one (k_corr_shapes + k_epilogue2):
//   k_wcorr  : the correlations, data window of KT + W - 1 samples in registers, vectors broadcast from shared
//              memory as 128-bit loads, taps unrolled at compile time per support class (whole window / rise half /
//              decay half: a flare shape touches only its half)            -> Yw[row][sample]
//   k_wsolve : one sample per thread -- Cholesky + L^-1, per shape solve, per template term, the grid sums in the
//              linear domain (log-domain redo for samples with a flare), the hypotheses folded into the log odds.
//              The y vectors of the A shapes stay in registers; the shapes' sums are prefetched four shapes ahead
//              with cp.async.  About 1 KB of state per sample caps the SM at 8 warps, so the kernel lives on
//              instruction-level parallelism: five templates per basic block, no run-time branch inside a block.
// Yw makes one round trip through HBM (8 (P(P+3)/2 + (P+2) S') bytes per sample each way, S' non-empty shapes).
// The one-kernel form k_window_wsep sits at 45 % of the FP64 pipe with 8 warps per SM and a tap loop of 113
// instructions for 64 DFMAs; a two-threads-per-sample variant of it was measured no faster (round 2).
#include "bfl_kernels.cuh"

#include <stdlib.h>
#include <algorithm>
#include <type_traits>

#include "bfl_devmath.cuh"

namespace bfl {

// --------------------------------------------------------------------------------------
// k_wcorr
// --------------------------------------------------------------------------------------
struct WCorrArgs {
  const double* dw; const double* uu;      // d/v and v of the chunk's curves, [B][seglen]
  int64_t seglen, N, seg0, k0, k1;
  int B, tiles_per_curve;
  int64_t total_wtiles;
  const double* tab; const int* off; const int* row;
  int n[2][3];
  int tab_len, nj;
  double* Y;                               // [rows][BK]
  int64_t BK;
  unsigned long long* tile_counter;        // dynamic warp tiles (as k_corr_shapes): the last warp to leave re-arms both
  unsigned long long* done_counter;
};

constexpr int WC_TPB = 128, WC_KT = 4;

// z[j] = sum over taps T0 .. T0+NT-1 of t[w - T0] * win[w + j]; `t` points at the vector's first stored tap
template <int T0, int NT, int KT, int WL>
__device__ __forceinline__ void corr_taps(const double* __restrict__ t, const double (&win)[WL], double (&z)[KT]) {
  constexpr int NPAIR = NT / 2;
  double acc[KT][2];
#pragma unroll
  for (int j = 0; j < KT; j++) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
  const double2* t2 = reinterpret_cast<const double2*>(t);
#pragma unroll
  for (int w2 = 0; w2 < NPAIR; w2++) {
    const double2 m = t2[w2];
#pragma unroll
    for (int j = 0; j < KT; j++) {
      acc[j][w2 & 1] = fma(m.x, win[T0 + 2 * w2 + j], acc[j][w2 & 1]);
      acc[j][w2 & 1] = fma(m.y, win[T0 + 2 * w2 + 1 + j], acc[j][w2 & 1]);
    }
  }
  if constexpr ((NT & 1) != 0) {
    const double ml = t[NT - 1];
#pragma unroll
    for (int j = 0; j < KT; j++) z[j] = fma(ml, win[T0 + NT - 1 + j], acc[j][0]) + acc[j][1];
  } else {
#pragma unroll
    for (int j = 0; j < KT; j++) z[j] = acc[j][0] + acc[j][1];
  }
}

// W is a compile-time window length: instantiated for 21, 33 and 55 (the reference's default and the shorter lengths its
// scripts use); other lengths keep k_window_wsep / k_window_weighted
template <int W, int KT>
__global__ void __launch_bounds__(WC_TPB, 3) k_wcorr(const WCorrArgs a) {
  extern __shared__ double smem[];
  constexpr int H = W / 2, TKW = 32 * KT;
  constexpr int T0H = H & ~1, NTL = H + 1, NTH = W - T0H;
  constexpr int wstride = (TKW + 2 * H + 2) & ~1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* swin = smem + warp * 2 * wstride;                 // [2][wstride]: 1/v, d/v
  double* stab = smem + (WC_TPB / 32) * 2 * wstride;        // packed vectors
  int* soff = reinterpret_cast<int*>(stab + a.tab_len);     // [nj]
  int* srow = soff + a.nj;                                  // [nj]
  {
    const double2* src = reinterpret_cast<const double2*>(a.tab);
    double2* dst = reinterpret_cast<double2*>(stab);
    for (int i = tid; i < a.tab_len / 2; i += WC_TPB) dst[i] = __ldg(src + i);
    for (int i = tid; i < a.nj; i += WC_TPB) { soff[i] = a.off[i]; srow[i] = a.row[i]; }
  }
  __syncthreads();
  const int64_t nk = a.k1 - a.k0;
  for (;;) {
    long long wt = 0;
    if (lane == 0) wt = (long long)atomicAdd(a.tile_counter, 1ULL);
    wt = __shfl_sync(0xffffffffu, wt, 0);
    if (wt >= a.total_wtiles) {
      if (lane == 0 && atomicAdd(a.done_counter, 1ULL) == (unsigned long long)gridDim.x * (WC_TPB / 32) - 1ULL) {
        *a.tile_counter = 0ULL;
        *a.done_counter = 0ULL;
      }
      break;
    }
    const int tile = (int)(wt % a.tiles_per_curve);
    const int b = (int)(wt / a.tiles_per_curve);
    const int64_t kbase = a.k0 + (int64_t)tile * TKW;
    {
      const double* du = a.uu + (size_t)b * a.seglen;
      const double* dd = a.dw + (size_t)b * a.seglen;
      const int64_t lo = a.seg0, hi = min(a.N, a.seg0 + a.seglen);
      __syncwarp();
      for (int i = lane; i < TKW + 2 * H; i += 32) {
        const int64_t sidx = kbase - H + i;
        const bool in = sidx >= lo && sidx < hi;
        swin[i] = in ? 1.0 / __ldg(du + (sidx - a.seg0)) : 0.0;
        swin[wstride + i] = in ? __ldg(dd + (sidx - a.seg0)) : 0.0;
      }
      __syncwarp();
    }
    const int64_t kfirst = kbase + (int64_t)lane * KT;
    bool valid[KT];
#pragma unroll
    for (int j = 0; j < KT; j++) {
      const int64_t k = kfirst + j;
      valid[j] = (k < a.k1) && (k >= H) && (k < a.N - H);
    }
    double* yout = a.Y + (size_t)b * nk + (kfirst - a.k0);
    int jb = 0;
#pragma unroll 1
    for (int st = 0; st < 2; st++) {
      double win[W + KT - 1];
#pragma unroll
      for (int e = 0; e < W + KT - 1; e++) win[e] = swin[st * wstride + lane * KT + e];
      auto put = [&](int j, const double (&y)[KT]) {
        double* o = yout + (size_t)srow[j] * a.BK;
#pragma unroll
        for (int q = 0; q < KT; q++)
          if (valid[q]) o[q] = y[q];
      };
#pragma unroll 1
      for (int i = 0; i < a.n[st][0]; i++, jb++) {
        double y[KT];
        corr_taps<0, W, KT, W + KT - 1>(stab + soff[jb], win, y);
        put(jb, y);
      }
#pragma unroll 1
      for (int i = 0; i < a.n[st][1]; i++, jb++) {
        double y[KT];
        corr_taps<0, NTL, KT, W + KT - 1>(stab + soff[jb], win, y);
        put(jb, y);
      }
#pragma unroll 1
      for (int i = 0; i < a.n[st][2]; i++, jb++) {
        double y[KT];
        corr_taps<T0H, NTH, KT, W + KT - 1>(stab + soff[jb], win, y);
        put(jb, y);
      }
    }
  }
}

// --------------------------------------------------------------------------------------
// k_wsolve
// --------------------------------------------------------------------------------------
struct WSolveArgs {
  const double* Y;                         // [rows][BK] of this chunk
  int64_t BK, nk, N, k0;
  int64_t ki0, part_stride;                // chunk's first sample in the whole batch's [B * nk]; B * nk
  int64_t total_wtiles;                    // warp tiles of 16 samples
  const int* shrow;
  const double* lp; const double* lw; const short* idxA; const int* sh_g0; const int* sh_g1;
  int H, G, S, nhyp;
  int hyp_begin[MAXH + 1];
  int hyp_half[MAXH];
  int hyp_A0[MAXH], hyp_nA[MAXH], hyp_B0[MAXH], hyp_nB[MAXH];
  double logdetG;
  double2* part;
  double* bgabs;
  double* lnO;                             // not null: combine the hypotheses here (find.py:853-862) and write the log odds
  int noisepoly;
  unsigned long long* tile_counter;        // next warp tile: tiles with a flare cost a second pass, a static split leaves
  unsigned long long* done_counter;        // SMs idle (ncu: 23 % of the launch); the last warp to leave re-arms both
};

constexpr int WS_TPB = 128, WS_MINB = 2;   // one sample per thread, two CTAs (8 warps) per SM: 255 registers, no spills
constexpr int WS_KD = 4;                   // shapes whose window sums are in flight per warp (cp.async ring)
constexpr int WS_NA = 10;                  // A shapes per separable hypothesis the kernel keeps in registers (even)

// 1/sqrt(s): MUFU.RSQ64H + one third-order correction (relative error ~1e-18 from a 2^-20 seed); NaN for s < 0
__device__ __forceinline__ double rsqrt_seeded(double s) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(s));
  const double g = s * y0, hy = 0.5 * y0;
  const double r = fma(-g, hy, 0.5);               // (1 - s y0^2) / 2
  return fma(y0, r * fma(1.5, r, 1.0), y0);
}

template <int N>
__device__ __forceinline__ void lse_merge(double& mx, double& sm, const double (&x)[N], const double (&w)[N]) {
  double bm = x[0];
#pragma unroll
  for (int c = 1; c < N; c++) bm = fmax(bm, x[c]);
  double bs = 0.0;
#pragma unroll
  for (int c = 0; c < N; c++) bs = fma(w[c], exp_nonpos_poly(x[c] - bm), bs);
  const double d = bm - mx;
  const double e = exp_nonpos_poly(-fabs(d));
  const bool up = __double2hiint(d) >= 0;
  sm = up ? fma(sm, e, bs) : fma(bs, e, sm);
  mx = up ? bm : mx;
}

// One sample per thread.  The y vectors of a separable hypothesis' A shapes (NA x P doubles) stay in REGISTERS, so a
// template's cross term yA.yB costs P FMAs and no shared-memory traffic; the loop over A shapes is unrolled at
// compile time and a per-(B shape, A shape) weight table (zero for pairs the prior excludes) replaces index lists.
// Two evaluation strategies, as in k_epilogue2:
//  * LINEAR (first): nearly every sample is noise, every |z_g| a few units.  The grid sum is accumulated directly,
//    S_h = sum_g c_g w_g E(z_g) with c_g = exp(lp_g + lw_g - ref_h) fixed per plan, w_g = 1/sqrt(s_g) and
//    E = exp(phi) from the 8-byte table (exphi2_one); full-range hypotheses use exp(z^2/2 - ZC^2/2).  No range
//    reduction, no running maximum.
//  * LOG (fallback): a thread that met |z| > ZC (a flare or a deep dip), a non-positive s or a non-finite sum redoes
//    its sample with the phi table and an online log-sum-exp, any z.
template <int N>
__device__ __forceinline__ void cp_async_wait_pending() { asm volatile("cp.async.wait_group %0;" : : "n"(N) : "memory"); }

template <int P, int NA, int TPB, int MINB, int KD>
__global__ void __launch_bounds__(TPB, MINB) k_wsolve(const WSolveArgs a) {
  static_assert((KD & (KD - 1)) == 0, "ring depth is a power of two");
  static_assert(NA % 2 == 0, "A shapes are evaluated two at a time");
  extern __shared__ double smem[];
  constexpr int NV = P + 2;                        // per shape: y[P] (f[P] before the solve), s' (e), r' (r0)
  constexpr int NLI = P * (P + 1) / 2, NL = NLI + P;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr double ZC = BFL_EXPHI2_ZC;
  const int tid = threadIdx.x, lane = tid & 31;
  double* sT = smem;                               // [NI] E table
  double* sC = sT + ((BFL_EXPHI2_NI + 1) & ~1);    // [S][NA] weight of template (B shape s, A shape i), 0 if absent;
                                                   //         dense shapes: [s][0] = weight of their one template
  double* sK = sC + (size_t)a.S * NA;              // [G] lp + lw (log-domain pass)
  double* sRef = sK + ((a.G + 1) & ~1);            // [MAXH] reference value of a hypothesis' linear sum
  double* sL = sRef + MAXH;                        // [NL][TPB] per-thread L^-1 (packed) and yc
  double* sAB = sL + (size_t)NL * TPB;             // [2 NA][TPB] per-thread s'_A, r'_A
  double* sRel = sAB + (size_t)2 * NA * TPB;       // [MAXH][TPB] per-thread log marginal of every hypothesis (fused combine)
  double* sRing = sRel + (size_t)MAXH * TPB;       // [TPB / 32][KD][NV][32] prefetched window sums
  int* sShRow = reinterpret_cast<int*>(sRing + (size_t)(TPB / 32) * KD * NV * 32);   // [S]
  int* sGi = sShRow + a.S;                         // [S][NA] template index of (s, i) or -1
  int* sSeq = sGi + (size_t)a.S * NA;              // [S + 1] Yw rows of the shapes in the order a sample consumes them (-1: empty
                                                   //         shape); [S] = how many
  for (int i = tid; i < BFL_EXPHI2_NI; i += TPB) sT[i] = BFL_EXPHI2_TAB[i];
  for (int i = tid; i < a.G; i += TPB) sK[i] = a.lp[i] + a.lw[i];
  for (int i = tid; i < a.S; i += TPB) sShRow[i] = a.shrow[i];
  for (int i = tid; i < a.S * NA; i += TPB) { sC[i] = 0.0; sGi[i] = -1; }
  if (tid == 0) {
    int n = 0;
    for (int h = 0; h < a.nhyp; h++) {
      if (a.hyp_begin[h] >= a.hyp_begin[h + 1]) continue;
      for (int i = 0; i < a.hyp_nA[h]; i++) sSeq[n++] = a.shrow[a.hyp_A0[h] + i];
      for (int j = 0; j < a.hyp_nB[h]; j++) sSeq[n++] = a.shrow[a.hyp_B0[h] + j];
    }
    sSeq[a.S] = n;
  }
  __syncthreads();
  if (tid < a.nhyp) {
    double m = neg_inf();
    for (int g = a.hyp_begin[tid]; g < a.hyp_begin[tid + 1]; g++) m = fmax(m, sK[g]);
    sRef[tid] = (m > neg_inf()) ? m : 0.0;
  }
  __syncthreads();
  for (int sh = tid; sh < a.S; sh += TPB) {
    for (int g = a.sh_g0[sh]; g < a.sh_g1[sh]; g++) {
      int h = 0;
      while (h + 1 < a.nhyp && g >= a.hyp_begin[h + 1]) h++;
      const int i = max((int)a.idxA[g], 0);        // dense templates (idxA = -1) use column 0
      if (i < NA) { sC[sh * NA + i] = exp(sK[g] - sRef[h]); sGi[sh * NA + i] = g; }
    }
  }
  __syncthreads();
  const unsigned aT = (unsigned)__cvta_generic_to_shared(sT);
  double* myL = sL + tid;
  double* myAB = sAB + tid;
  double* myRel = sRel + tid;
  const int64_t BK = a.BK;
  const int nseq = sSeq[a.S];
  const unsigned ring = (unsigned)__cvta_generic_to_shared(sRing + (size_t)(tid >> 5) * KD * NV * 32 + lane);
  constexpr unsigned RQ = 32 * 8, RS = NV * RQ;    // bytes between the values of a shape, between ring slots

  for (;;) {
  long long wt = 0;
  if (lane == 0) wt = (long long)atomicAdd(a.tile_counter, 1ULL);
  wt = __shfl_sync(FULL, wt, 0);
  if (wt >= a.total_wtiles) {
    if (lane == 0 && atomicAdd(a.done_counter, 1ULL) == (unsigned long long)gridDim.x * (TPB / 32) - 1ULL) {
      *a.tile_counter = 0ULL;
      *a.done_counter = 0ULL;
    }
    break;
  }
  const int64_t kif = wt * 32 + lane;
  bool active = kif < BK;
  {
    const int64_t k = a.k0 + kif % a.nk;
    active = active && k >= a.H && k < a.N - a.H;           // interior samples only (k_wcorr wrote no others)
  }
  const unsigned bal = __ballot_sync(FULL, active);
  if (bal == 0u) continue;
  // idle lanes shadow the warp's first interior sample, so every lane runs the same code on defined sums
  const int64_t kfirst_active = __shfl_sync(FULL, kif, __ffs(bal) - 1);   // (every lane takes part in the shuffle)
  const int64_t kic = active ? kif : kfirst_active;
  const double* Yk = a.Y + kic;

  // ---- A = Q U Q^T = L L^T, yc = L^-1 Q U d, then L^-1 explicitly; both go to the thread's shared-memory slice ----
  double logdetA = 0.0, quad = 0.0;
  {
    double A[P][P], c[P], yc[P], piv[P];
    {
      const double* yp = Yk;
#pragma unroll
      for (int p = 0; p < P; p++) {
#pragma unroll
        for (int q = p; q < P; q++, yp += BK) A[p][q] = *yp;
      }
#pragma unroll
      for (int p = 0; p < P; p++, yp += BK) c[p] = *yp;
    }
    double L[P][P], inv[P], Li[NLI];
#pragma unroll
    for (int i = 0; i < P; i++) {
      double d = A[i][i];
#pragma unroll
      for (int t = 0; t < i; t++) d = fma(-L[i][t], L[i][t], d);
      piv[i] = d;
      const double r = rsqrt_seeded(d);
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
#pragma unroll
    for (int i = 0; i < P; i++) {
      Li[i * (i + 1) / 2 + i] = inv[i];
#pragma unroll
      for (int j = 0; j < i; j++) {
        double x = 0.0;
#pragma unroll
        for (int t = j; t < i; t++) x = fma(L[i][t], Li[t * (t + 1) / 2 + j], x);
        Li[i * (i + 1) / 2 + j] = -x * inv[i];
      }
    }
    {
      // log det A = sum of the log pivots: ONE log of their product unless that leaves the normal range
      double prod = piv[0];
#pragma unroll
      for (int i = 1; i < P; i++) prod *= piv[i];
      if (prod > 1.0e-280 && prod < 1.0e280) {
        logdetA = log(prod);
      } else {
#pragma unroll
        for (int i = 0; i < P; i++) logdetA += log(piv[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < NLI; i++) myL[i * TPB] = Li[i];
#pragma unroll
    for (int i = 0; i < P; i++) myL[(NLI + i) * TPB] = yc[i];
  }
  double lnBbg = -0.5 * a.logdetG - 0.5 * logdetA + 0.5 * quad + 0.5 * P * (BFL_LN2 + BFL_LNPI);
  const bool bad = !isfinite(lnBbg);               // non-finite data or variances, log_marg_amp_full.c:20,25
  if (bad) lnBbg = neg_inf();
  const int64_t ko = a.ki0 + kif;                  // the sample in the whole batch
  if (a.bgabs && active) a.bgabs[ko] = lnBbg;

  // window sums of shape sh from Yw (zero for an empty shape)
  auto load_shape = [&](int sh, double (&F)[NV]) {
    const int row = sShRow[sh];
    if (row >= 0) {
      const double* yp = Yk + (size_t)row * BK;
#pragma unroll
      for (int q = 0; q < NV; q++, yp += BK) F[q] = *yp;
    } else {
#pragma unroll
      for (int q = 0; q < NV; q++) F[q] = 0.0;
    }
  };
  // The LINEAR pass reads the shapes in the fixed order sSeq; their sums are fetched KD shapes ahead with cp.async
  // (LDGSTS, no registers) into the warp's ring: one group per shape, committed even when nothing is fetched so that
  // "all but the KD - 1 newest groups are complete" always means "shape n has arrived".
  auto issue = [&](int n) {
    if (n < nseq) {
      const int row = sSeq[n];
      if (row >= 0) {
        const double* yp = Yk + (size_t)row * BK;
        const unsigned dst = ring + (unsigned)(n & (KD - 1)) * RS;
#pragma unroll
        for (int q = 0; q < NV; q++, yp += BK) cp_async_f64(dst + q * RQ, yp);
      }
    }
    cp_async_commit();
  };
  auto take = [&](int n, double (&F)[NV]) {
    cp_async_wait_pending<KD - 1>();
    const unsigned src = ring + (unsigned)(n & (KD - 1)) * RS;
    if (sSeq[n] >= 0) {                            // (warp-uniform)
#pragma unroll
      for (int q = 0; q < NV; q++) F[q] = lds_f64(src + q * RQ);
    } else {
#pragma unroll
      for (int q = 0; q < NV; q++) F[q] = 0.0;
    }
    issue(n + KD);
  };
  // (f, e, r0) -> (y = L^-1 f, s' = e - |y|^2, r' = r0 - y.yc)
  auto solve = [&](const double (&F)[NV], double (&V)[NV]) {
    double s = F[P], r = F[P + 1];
#pragma unroll
    for (int i = 0; i < P; i++) {
      double y = myL[(i * (i + 1) / 2) * TPB] * F[0];
#pragma unroll
      for (int q = 1; q <= i; q++) y = fma(myL[(i * (i + 1) / 2 + q) * TPB], F[q], y);
      V[i] = y;
    }
#pragma unroll
    for (int i = 0; i < P; i++) { s = fma(-V[i], V[i], s); r = fma(-V[i], myL[(NLI + i) * TPB], r); }
    V[P] = s;
    V[P + 1] = r;
  };

  // two shapes at once: the factor is read once and the two dependency chains interleave
  auto solve2 = [&](const double (&F0)[NV], const double (&F1)[NV], double (&V0)[NV], double (&V1)[NV]) {
    double s0 = F0[P], r0 = F0[P + 1], s1 = F1[P], r1 = F1[P + 1];
#pragma unroll
    for (int i = 0; i < P; i++) {
      const double l0 = myL[(i * (i + 1) / 2) * TPB];
      double y0 = l0 * F0[0], y1 = l0 * F1[0];
#pragma unroll
      for (int q = 1; q <= i; q++) {
        const double l = myL[(i * (i + 1) / 2 + q) * TPB];
        y0 = fma(l, F0[q], y0);
        y1 = fma(l, F1[q], y1);
      }
      V0[i] = y0;
      V1[i] = y1;
    }
#pragma unroll
    for (int i = 0; i < P; i++) {
      const double c = myL[(NLI + i) * TPB];
      s0 = fma(-V0[i], V0[i], s0); r0 = fma(-V0[i], c, r0);
      s1 = fma(-V1[i], V1[i], s1); r1 = fma(-V1[i], c, r1);
    }
    V0[P] = s0; V0[P + 1] = r0;
    V1[P] = s1; V1[P + 1] = r1;
  };

  // One pass over the hypotheses.  LIN: linear-domain sums, returns true when this sample has to be redone.
  auto hypotheses = [&](auto lin_tag) -> bool {
    constexpr bool LIN = decltype(lin_tag)::value;
    bool big = false;
    int nsh = 0;                                    // LIN: position in sSeq
    if constexpr (LIN) {
#pragma unroll
      for (int n = 0; n < KD; n++) issue(n);
    }
    for (int h = 0; h < a.nhyp; h++) {
      if (a.hyp_begin[h] >= a.hyp_begin[h + 1]) continue;
      const bool half = a.hyp_half[h] != 0;
      const double hc = half ? 0.5 * (BFL_LNPI - BFL_LN2) : 0.5 * (BFL_LN2 + BFL_LNPI);
      const int nA = a.hyp_nA[h], A0 = a.hyp_A0[h], B0 = a.hyp_B0[h], nB = a.hyp_nB[h];
      double S = 0.0, zz = 0.0;                     // LIN: the sum, the largest z^2 of a full-range hypothesis
      unsigned kmax = 0;                            // LIN: the largest table index (half-range)
      double mx = LSE_EMPTY, sm = 0.0;              // LOG
      // HALF is a compile-time tag: with a run-time branch inside, every term is its own basic block and the K
      // dependency chains of a group are scheduled one after the other instead of interleaved
      auto lin_term = [&](auto half_tag, double s, double r, double c) {
        const double w = rsqrt_seeded(s);
        const double z = r * w;
        if constexpr (decltype(half_tag)::value) {
          S = fma(c * w, exphi2_one(z + ZC, aT, kmax), S);
        } else {
          const double q = z * z;
          zz = fmax(zz, q);
          S = fma(c * w, exp_nonpos_poly(fma(0.5, q, -0.5 * ZC * ZC)), S);
        }
      };
      auto log_term = [&](double s, double r, double kg, bool valid, double& x, double& w) {
        w = rsqrt(s);
        const double z = r * w;
        double val;
        if (half) {
          val = phi_tab_rows(z, BFL_PHI_TAB);
          if (z < BFL_PHI_ZFAR) val = phi_half_far(z);
        } else {
          val = 0.5 * z * z;
        }
        x = valid ? val + kg : LSE_EMPTY;
        w = valid ? w : 0.0;
      };
      // K templates at once (independent dependency chains: the kernel runs at 8 warps per SM and lives on ILP)
      auto many = [&](auto ktag, const double* sv, const double* rv, const double* cw, const int* gi) {
        constexpr int K = decltype(ktag)::value;
        if constexpr (LIN) {
          if (half) {
#pragma unroll
            for (int c = 0; c < K; c++) lin_term(std::true_type{}, sv[c], rv[c], cw[c]);
          } else {
#pragma unroll
            for (int c = 0; c < K; c++) lin_term(std::false_type{}, sv[c], rv[c], cw[c]);
          }
        } else {
          double x[K], w[K];
#pragma unroll
          for (int c = 0; c < K; c++) log_term(sv[c], rv[c], sK[max(gi[c], 0)], gi[c] >= 0, x[c], w[c]);
          lse_merge<K>(mx, sm, x, w);
        }
      };
      if (nA > 0) {
        // separable hypothesis: A shapes reduced once (y in registers, s' and r' in the slice), then every B shape
        // pairs with the A shapes its row of the weight table admits
        double yA[NA][P];
#pragma unroll
        for (int i = 0; i < NA; i += 2) {
          double F0[NV], F1[NV], V0[NV], V1[NV];
#pragma unroll
          for (int q = 0; q < NV; q++) { V0[q] = 0.0; V1[q] = 0.0; }
          if (i + 1 < nA) {
            if constexpr (LIN) { take(nsh++, F0); take(nsh++, F1); }
            else { load_shape(A0 + i, F0); load_shape(A0 + i + 1, F1); }
            solve2(F0, F1, V0, V1);
          } else if (i < nA) {
            if constexpr (LIN) take(nsh++, F0);
            else load_shape(A0 + i, F0);
            solve(F0, V0);
          }
#pragma unroll
          for (int q = 0; q < P; q++) { yA[i][q] = V0[q]; yA[i + 1][q] = V1[q]; }
          myAB[(2 * i) * TPB] = V0[P];
          myAB[(2 * i + 1) * TPB] = V0[P + 1];
          myAB[(2 * i + 2) * TPB] = V1[P];
          myAB[(2 * i + 3) * TPB] = V1[P + 1];
        }
        if constexpr (LIN) {
          // Software pipeline over the B shapes: the NEXT shape's solve (a serial chain of ~45 instructions) sits in
          // the same basic block as this shape's first five templates, so ptxas interleaves it with their chains.
          // `half` is hoisted out of the loop for the same reason (a branch ends the block).
          auto b_loop = [&](auto half_tag) {
            double Vc[NV];
            {
              double F[NV];
              take(nsh++, F);
              solve(F, Vc);
            }
            for (int j = 0; j < nB; j++) {
              double Fn[NV], Vn[NV];
              if (j + 1 < nB) {
                take(nsh++, Fn);
              } else {
#pragma unroll
                for (int q = 0; q < NV; q++) Fn[q] = 0.0;
              }
              const double* cw = sC + (size_t)(B0 + j) * NA;
              const int* gi = sGi + (size_t)(B0 + j) * NA;
              auto five = [&](auto itag) {
                constexpr int I0 = decltype(itag)::value;
                double sv[5], rv[5];
#pragma unroll
                for (int c = 0; c < 5; c++) {
                  double d = 0.0;
#pragma unroll
                  for (int q = 0; q < P; q++) d = fma(yA[I0 + c][q], Vc[q], d);
                  sv[c] = fma(-2.0, d, Vc[P] + myAB[(2 * (I0 + c)) * TPB]);
                  rv[c] = Vc[P + 1] + myAB[(2 * (I0 + c) + 1) * TPB];
                }
#pragma unroll
                for (int c = 0; c < 5; c++) lin_term(half_tag, sv[c], rv[c], cw[I0 + c]);   // weight 0: pair excluded
              };
              solve(Fn, Vn);
              five(std::integral_constant<int, 0>{});
              bool any = false;
#pragma unroll
              for (int c = 5; c < NA; c++) any = any || gi[c] >= 0;
              if (any) five(std::integral_constant<int, 5>{});   // (warp-uniform)
#pragma unroll
              for (int q = 0; q < NV; q++) Vc[q] = Vn[q];
            }
          };
          static_assert(NA == 10, "the grouping is written for ten A shapes");
          if (nB > 0) {
            if (half) b_loop(std::true_type{});
            else b_loop(std::false_type{});
          }
        } else {
        for (int j = 0; j < nB; j++) {
          double F[NV], V[NV];
          load_shape(B0 + j, F);
          solve(F, V);
          const double* cw = sC + (size_t)(B0 + j) * NA;
          const int* gi = sGi + (size_t)(B0 + j) * NA;
          // A shapes five at a time: five independent dependency chains per basic block
          auto group = [&](auto itag, auto ktag) {
            constexpr int I0 = decltype(itag)::value, K = decltype(ktag)::value;
            bool any = false;
#pragma unroll
            for (int c = 0; c < K; c++) any = any || gi[I0 + c] >= 0;
            if (!any) return;                                    // warp-uniform: the prior excludes these pairs
            double sv[K], rv[K];
#pragma unroll
            for (int c = 0; c < K; c++) {
              double d = 0.0;
#pragma unroll
              for (int q = 0; q < P; q++) d = fma(yA[I0 + c][q], V[q], d);
              sv[c] = fma(-2.0, d, V[P] + myAB[(2 * (I0 + c)) * TPB]);
              rv[c] = V[P + 1] + myAB[(2 * (I0 + c) + 1) * TPB];
            }
            many(ktag, sv, rv, cw + I0, gi + I0);
          };
          group(std::integral_constant<int, 0>{}, std::integral_constant<int, 5>{});
          group(std::integral_constant<int, 5>{}, std::integral_constant<int, 5>{});
        }
        }
      } else {
        // dense hypothesis: one template per shape, two shapes at a time
        for (int j = 0; j < nB; j += 2) {
          const int j1 = min(j + 1, nB - 1);
          double F0[NV], F1[NV], V0[NV], V1[NV];
          if constexpr (LIN) {
            take(nsh++, F0);
            if (j + 1 < nB) {
              take(nsh++, F1);
            } else {
#pragma unroll
              for (int q = 0; q < NV; q++) F1[q] = F0[q];
            }
          } else {
            load_shape(B0 + j, F0);
            load_shape(B0 + j1, F1);
          }
          solve2(F0, F1, V0, V1);
          const double cw[2] = {sC[(size_t)(B0 + j) * NA], (j + 1 < nB) ? sC[(size_t)(B0 + j1) * NA] : 0.0};
          const int gi[2] = {sGi[(size_t)(B0 + j) * NA], (j + 1 < nB) ? sGi[(size_t)(B0 + j1) * NA] : -1};
          const double sv[2] = {V0[P], V1[P]}, rv[2] = {V0[P + 1], V1[P + 1]};
          many(std::integral_constant<int, 2>{}, sv, rv, cw, gi);
        }
      }
      if constexpr (LIN) {
        big = big || (!bad && (kmax > (unsigned)(BFL_EXPHI2_NI - 1) || zz > ZC * ZC || !(S > 0.0 && S < 1.0e300)));
        double m = sRef[h] + hc + (half ? 0.0 : 0.5 * ZC * ZC);
        if (bad) { m = lnBbg - lnBbg; S = m; }       // NaN, as (-inf) - (-inf) in the general path
        if (a.lnO) myRel[h * TPB] = (m != m || S != S) ? m + S : ((S > 0.0) ? m + log(S) : neg_inf());   // as k_combine
        else if (active) a.part[(size_t)h * a.part_stride + ko] = make_double2(m, S);
      } else {
        if (bad) { mx = lnBbg - lnBbg; sm = mx; }
        if (a.lnO) myRel[h * TPB] = (mx != mx || sm != sm) ? mx + sm : ((sm > 0.0) ? (mx + hc) + log(sm) : neg_inf());
        else if (active) a.part[(size_t)h * a.part_stride + ko] = make_double2(mx + hc, sm);
      }
    }
    return big;
  };
  if (hypotheses(std::true_type{})) hypotheses(std::false_type{});
  if (a.lnO && active) {
    // the reference's fold (find.py:853-862; the same arithmetic as k_combine): numerator = the signal hypothesis,
    // denominator = logplus over [polynomial-only (= 0 in these units), noise models]
    const bool h0 = a.hyp_begin[0] < a.hyp_begin[1];
    double lnO = h0 ? myRel[0] : neg_inf();
    const int nnoise = a.nhyp - 1 + (a.noisepoly ? 1 : 0);
    if (nnoise > 0) {
      double denom = neg_inf();
      if (a.noisepoly) denom = logplus_dev(denom, 0.0);
      for (int h = 1; h < a.nhyp; h++)
        denom = logplus_dev(denom, (a.hyp_begin[h] < a.hyp_begin[h + 1]) ? myRel[h * TPB] : neg_inf());
      lnO = lnO - denom;
    } else {
      lnO = lnO + lnBbg;                            // no noise models: signal versus Gaussian noise
    }
    a.lnO[ko] = lnO;
  }
  }   // warp-tile loop
}

// --------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------
bool wsplit_window_supported(int W) { return W == 21 || W == 33 || W == 55; }

static size_t wcorr_smem(const PlanView& pv) {
  const int wstride = (32 * WC_KT + 2 * (pv.W / 2) + 2) & ~1;
  return sizeof(double) * ((size_t)(WC_TPB / 32) * 2 * wstride + (size_t)pv.wc_tab_len) + sizeof(int) * (size_t)(2 * pv.wc_nj + 2);
}

static size_t wsolve_smem(const PlanView& pv) {
  const int NL = pv.P * (pv.P + 1) / 2 + pv.P;
  return sizeof(double) * ((size_t)((BFL_EXPHI2_NI + 1) & ~1) + (size_t)pv.S * WS_NA + (size_t)((pv.G + 1) & ~1) + MAXH +
                           (size_t)(NL + 2 * WS_NA + MAXH) * WS_TPB + (size_t)(WS_TPB / 32) * WS_KD * (pv.P + 2) * 32) +
         sizeof(int) * ((size_t)pv.S * (WS_NA + 1) + (size_t)((pv.S + 2) & ~1));
}

bool wsplit_fits(const PlanView& pv) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("BFL_NO_WSPLIT"); off = (e && atoi(e)) ? 1 : 0; }
  if (off || pv.wc_tab == nullptr || !wsplit_window_supported(pv.W) || pv.P < 1 || pv.P > 6 || (pv.wc_tab_len % 2) != 0) return false;
  for (int h = 0; h < pv.nhyp; h++)
    if (pv.hyp_nA[h] > WS_NA) return false;         // the A shapes' y vectors live in registers
  return wcorr_smem(pv) <= 72 * 1024 && wsolve_smem(pv) <= 112 * 1024;
}

// Yw for a chunk of curves: as many as fit `cap` bytes (2 GiB unless BFL_WSPLIT_CAP_MB says otherwise)
size_t wsplit_scratch_bytes(const PlanView& pv, int B, int64_t nk, int* chunk_curves) {
  static long long cap_mb = -1;
  if (cap_mb < 0) { const char* e = getenv("BFL_WSPLIT_CAP_MB"); cap_mb = (e && atoll(e) > 0) ? atoll(e) : 2048; }
  const size_t per_curve = sizeof(double) * (size_t)pv.wc_rows * (size_t)nk;
  const size_t cap = (size_t)cap_mb << 20;
  int ch = (int)std::min<size_t>((size_t)B, std::max<size_t>(1, cap / std::max<size_t>(per_curve, 1)));
  *chunk_curves = ch;
  return per_curve * (size_t)ch;
}

template <int P, int TPB, int MINB>
static cudaError_t launch_wsolve_cfg(const WSolveArgs& ws, size_t smem, int sms, cudaStream_t s) {
  cudaError_t e = cudaFuncSetAttribute(k_wsolve<P, WS_NA, TPB, MINB, WS_KD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int64_t ctas = (ws.total_wtiles + TPB / 32 - 1) / (TPB / 32);
  k_wsolve<P, WS_NA, TPB, MINB, WS_KD><<<(unsigned)std::min<int64_t>(ctas, (int64_t)MINB * sms), TPB, smem, s>>>(ws);
  return cudaGetLastError();
}

template <int P>
static cudaError_t launch_wsolve(const WSolveArgs& ws, size_t smem, int sms, cudaStream_t s) {
  return launch_wsolve_cfg<P, WS_TPB, WS_MINB>(ws, smem, sms, s);
}

cudaError_t launch_window_wsplit(const PlanView& pv, const WsepArgs& a, double* Yw, int chunk_curves, int sms,
                                 double* d_lnO, int noisepoly, unsigned long long* counters, cudaStream_t s,
                                 LaunchStats* st) {
  const int64_t nk = a.k1 - a.k0;
  const size_t smem_c = wcorr_smem(pv), smem_s = wsolve_smem(pv);
  cudaError_t e = cudaSuccess;
  switch (pv.W) {
    case 21: e = cudaFuncSetAttribute(k_wcorr<21, WC_KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c); break;
    case 33: e = cudaFuncSetAttribute(k_wcorr<33, WC_KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c); break;
    case 55: e = cudaFuncSetAttribute(k_wcorr<55, WC_KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c); break;
    default: return cudaErrorInvalidValue;
  }
  if (e != cudaSuccess) return e;
  // chunks of equal size (a short last chunk would leave the correlation kernel's last round of tiles half empty)
  const int nchunks = (a.B + chunk_curves - 1) / chunk_curves;
  const int chunk = (a.B + nchunks - 1) / nchunks;
  for (int b0 = 0; b0 < a.B; b0 += chunk) {
    const int Bc = std::min(chunk, a.B - b0);
    WCorrArgs wc;
    wc.dw = a.dw + (size_t)b0 * a.seglen; wc.uu = a.uu + (size_t)b0 * a.seglen;
    wc.seglen = a.seglen; wc.N = a.N; wc.seg0 = a.seg0; wc.k0 = a.k0; wc.k1 = a.k1;
    wc.B = Bc;
    wc.tiles_per_curve = (int)((nk + 32 * WC_KT - 1) / (32 * WC_KT));
    wc.total_wtiles = (int64_t)wc.tiles_per_curve * Bc;
    wc.tab = pv.wc_tab; wc.off = pv.wc_off; wc.row = pv.wc_row;
    for (int i = 0; i < 2; i++)
      for (int j = 0; j < 3; j++) wc.n[i][j] = pv.wc_n[i][j];
    wc.tab_len = pv.wc_tab_len; wc.nj = pv.wc_nj;
    wc.Y = Yw; wc.BK = (int64_t)Bc * nk;
    wc.tile_counter = counters + 1; wc.done_counter = counters + 3;
    const int64_t ctas_c = (wc.total_wtiles + WC_TPB / 32 - 1) / (WC_TPB / 32);
    const unsigned grid_c = (unsigned)std::min<int64_t>(ctas_c, (int64_t)3 * sms);
    if (pv.W == 21) k_wcorr<21, WC_KT><<<grid_c, WC_TPB, smem_c, s>>>(wc);
    else if (pv.W == 33) k_wcorr<33, WC_KT><<<grid_c, WC_TPB, smem_c, s>>>(wc);
    else k_wcorr<55, WC_KT><<<grid_c, WC_TPB, smem_c, s>>>(wc);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    st->launches++;

    WSolveArgs ws;
    ws.Y = Yw; ws.BK = wc.BK; ws.nk = nk; ws.N = a.N; ws.k0 = a.k0;
    ws.ki0 = (int64_t)b0 * nk; ws.part_stride = (int64_t)a.B * nk;
    ws.total_wtiles = (ws.BK + 31) / 32;
    ws.shrow = pv.wc_shrow; ws.lp = a.lp; ws.lw = a.lw; ws.idxA = a.idxA; ws.sh_g0 = a.sh_g0; ws.sh_g1 = a.sh_g1;
    ws.H = a.W / 2; ws.G = a.G; ws.S = a.S; ws.nhyp = a.nhyp;
    for (int h = 0; h <= a.nhyp; h++) ws.hyp_begin[h] = a.hyp_begin[h];
    for (int h = 0; h < a.nhyp; h++) {
      ws.hyp_half[h] = a.hyp_half[h];
      ws.hyp_A0[h] = a.hyp_A0[h]; ws.hyp_nA[h] = a.hyp_nA[h]; ws.hyp_B0[h] = a.hyp_B0[h]; ws.hyp_nB[h] = a.hyp_nB[h];
    }
    ws.logdetG = a.logdetG; ws.part = a.part; ws.bgabs = a.bgabs;
    ws.lnO = d_lnO; ws.noisepoly = noisepoly;
    ws.tile_counter = counters; ws.done_counter = counters + 2;
    switch (pv.P) {
      case 1: e = launch_wsolve<1>(ws, smem_s, sms, s); break;
      case 2: e = launch_wsolve<2>(ws, smem_s, sms, s); break;
      case 3: e = launch_wsolve<3>(ws, smem_s, sms, s); break;
      case 4: e = launch_wsolve<4>(ws, smem_s, sms, s); break;
      case 5: e = launch_wsolve<5>(ws, smem_s, sms, s); break;
      case 6: e = launch_wsolve<6>(ws, smem_s, sms, s); break;
      default: return cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return e;
    st->launches++;
  }
  return cudaSuccess;
}

}  // namespace bfl
