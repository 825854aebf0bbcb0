/* oracle/bfl_oracle.c -- CPU restatement of BayesFlare's sliding-window odds-ratio path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product package
 * (bayesflare_b200) never does and has no CPU fallback.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors,
 * so the pin is the reference itself: every function below is checked in
 * tests/test_oracle_vs_reference.py against the reference's own compiled code
 * (oracle/_ref, built from /root/reference by oracle/build_ref.py) and its Python
 * layers imported in the build container, and against tests/golden/ fixtures that were
 * generated from the reference (tests/golden/make_golden.py).
 *
 * Each function cites the reference lines it follows (paths relative to
 * /root/reference/bayesflare).  Floating-point evaluation ORDER follows the reference
 * where the reference's order is defined by its own source (the elimination, the
 * numpy pairwise summation used by np.sum); build with -ffp-contract=off.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_LNPI 1.14472988584940017414342735135
#define ORC_LN2 0.693147180559945309417232121458
#define ORC_MAXNM 16

/* ---- special function: log(erfc(x)) ----------------------------------------------
 * Stands in for GSL's gsl_sf_log_erfc (stats/log_marg_amp_full.c:63); same definition
 * as oracle/shim/gsl/gsl_sf_erf.h. */
double orc_log_erfc(double x)
{
  if (!(x > 25.0)) return log(erfc(x));
  double r = 1.0 / (2.0 * x * x), term = 1.0, sum = 1.0;
  for (int n = 1; n < 30; n++) {
    term *= -(2.0 * n - 1.0) * r;
    sum += term;
    if (fabs(term) < 1e-18 * fabs(sum)) break;
  }
  return -x * x - log(x) - 0.5 * ORC_LNPI + log(sum);
}

/* ---- log-space helpers: stats/general.pyx:519-547 (logplus), :576-607 (logtrapz) --- */
double orc_logplus(double x, double y)
{
  int bothinf = isinf(x) && isinf(y);
  if (bothinf && x < 0 && y < 0) return -INFINITY;
  if (x > y && !bothinf) return x + log(1 + exp(y - x));
  if (x <= y && !bothinf) return y + log(1 + exp(x - y));
  return INFINITY; /* both +inf, or a NaN: the reference's initial value of z */
}

double orc_logtrapz(const double *lx, long stride, const double *t, int n)
{
  double B = -INFINITY;
  for (int i = 0; i < n - 1; i++) {
    double z = orc_logplus(lx[i * stride], lx[(i + 1) * stride]);
    z = z + log(t[i + 1] - t[i]) - ORC_LN2;
    B = orc_logplus(z, B);
  }
  return B;
}

/* ---- numpy's summation order for np.sum over a contiguous float64 vector ----------
 * (NumPy pairwise_sum: <8 sequential; <=128 eight running partial sums combined as a
 * balanced tree, then the tail; >128 recursive halves rounded down to a multiple of 8).
 * The reference builds every windowed cross term with np.sum (stats/bayes.py:329,377,393),
 * so this order is part of the reference's arithmetic. */
double orc_pairwise_sum(const double *a, long n)
{
  if (n < 8) {
    double res = 0.;
    for (long i = 0; i < n; i++) res += a[i];
    return res;
  } else if (n <= 128) {
    double r[8];
    long i;
    for (int j = 0; j < 8; j++) r[j] = a[j];
    for (i = 8; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; j++) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res += a[i];
    return res;
  } else {
    long n2 = n / 2;
    n2 -= n2 % 8;
    return orc_pairwise_sum(a, n2) + orc_pairwise_sum(a + n2, n - n2);
  }
}

/* ---- analytic amplitude marginalisation: stats/log_marg_amp_full.c:3-68 -----------
 * M: Nm x Nm row-major, upper triangle used; D: Nm.  "Complete the square" one
 * amplitude at a time in natural order; sq[] are the successive pivots, c[][] the
 * linear/cross coefficients.  Products are evaluated (a*b)*inv as in the reference. */
double orc_marg_full(int nm_, const double *M, const double *D, double sigma, int lhr)
{
  const int n = nm_, last = n - 1;
  double sq[ORC_MAXNM], c[ORC_MAXNM][ORC_MAXNM];
  double Z = 0.;
  memset(c, 0, sizeof(c));
  for (int i = 0; i < n; i++) {
    sq[i] = M[i * n + i];
    c[i][i] = -2. * D[i];
    if (!isfinite(sq[i]) || !isfinite(c[i][i])) return -INFINITY;
    for (int j = i + 1; j < n; j++) {
      c[i][j] = 2. * M[i * n + j];
      if (!isfinite(c[i][j])) return -INFINITY;
    }
  }
  for (int i = 0; i < last; i++) {
    const double inv = 1. / sq[i], inv2 = 0.5 * inv, inv4 = 0.25 * inv;
    Z -= c[i][i] * c[i][i] * inv4;
    for (int k = i + 1; k < n; k++) sq[k] -= c[i][k] * c[i][k] * inv4;
    for (int j = i + 1; j < n; j++) {
      c[j][j] -= c[i][j] * c[i][i] * inv2;
      for (int k = j + 1; k < n; k++) c[j][k] -= c[i][j] * c[i][k] * inv2;
    }
  }
  const double X = sq[last], Y = c[last][last];
  double logL = 0.;
  for (int i = 0; i < n; i++) logL -= 0.5 * log(sq[i]);
  logL -= 0.5 * (Z - 0.25 * Y * Y / X) / (sigma * sigma);
  logL += n * log(sigma);
  const double log2pi = ORC_LN2 + ORC_LNPI, logpi_2 = ORC_LNPI - ORC_LN2;
  if (lhr == 1)
    logL += 0.5 * (double)last * log2pi + 0.5 * logpi_2 + orc_log_erfc(0.5 * Y / (sigma * sqrt(2. * X)));
  else
    logL += 0.5 * (double)n * log2pi;
  return logL;
}

/* ---- all amplitudes but the last marginalised: stats/log_marg_amp_full.c:71-134 ----
 * Replayed literally, including the Z-update guard at :107 that makes the result differ
 * from the exact Gaussian integral for Nm >= 5 (SURVEY.md note N3). */
double orc_marg_except_final(int nm_, const double *M, const double *D, double sigma)
{
  const int n = nm_, nm = n - 1, nmm = n - 2;
  double sq[ORC_MAXNM], c[ORC_MAXNM][ORC_MAXNM];
  double Y = 0., Z = 0.;
  memset(c, 0, sizeof(c));
  for (int i = 0; i < n; i++) {
    sq[i] = M[i * n + i];
    c[i][i] = -2. * D[i];
    if (!isfinite(sq[i]) || !isfinite(c[i][i])) return -INFINITY;
    for (int j = i + 1; j < n; j++) {
      c[i][j] = 2. * M[i * n + j];
      if (!isfinite(c[i][j])) return -INFINITY;
    }
  }
  for (int i = 0; i < nmm; i++) {
    const double inv = 1. / sq[i], inv2 = 0.5 * inv, inv4 = 0.25 * inv;
    for (int j = i; j < n; j++) {
      for (int k = i; k < n; k++) {
        /* the reference reads c[i][j] with j possibly < ... always j>=i here; entries
         * below the diagonal are never written and stay 0 */
        if ((j != i + 1 && j != nmm) && (k != i + 1 && k != nmm)) Z -= c[i][j] * c[i][k] * inv4;
        if (j == i) {
          if (k > j && k < nm) sq[k] -= c[j][k] * c[j][k] * inv4;
        } else {
          if (k == i && j < nm) c[j][j] -= c[i][j] * c[i][k] * inv2;
          else if (k > j) c[j][k] -= c[i][j] * c[i][k] * inv2;
        }
      }
    }
  }
  const double X = sq[nmm];
  for (int i = nmm; i < n; i++) Y += c[nmm][i];
  Z += (sq[nm] + c[nm][nm]);
  double logL = 0.;
  for (int i = 0; i < nm; i++) logL -= 0.5 * log(sq[i]);
  logL -= 0.5 * (Z - 0.25 * Y * Y / X) / (sigma * sigma);
  logL += nm * log(sigma);
  logL += 0.5 * (double)nm * (ORC_LN2 + ORC_LNPI);
  return logL;
}

/* ---- windowed cross terms: stats/bayes.py:315-329 (bgcross), :367-377 (mdcross),
 * :380-393 (mdbgcross).  out[k] = np.sum((a*b)[window] / v[window]) with the window
 * truncated at the series ends exactly as the reference slices it. */
void orc_window_cross(const double *a, const double *b, const double *v, long N, int W, double *out)
{
  const int h = W / 2;
  double *tmp = (double *)malloc(sizeof(double) * (size_t)W);
  for (long k = 0; k < N; k++) {
    int w0 = 0, w1 = W; /* template index range [w0,w1) that overlaps the data */
    if (k < h) w0 = (int)(h - k);
    if (k >= N - h) w1 = (int)(W - (k + h + 1 - N));
    const long base = k - h; /* data index of template element 0 */
    int n = 0;
    for (int w = w0; w < w1; w++) tmp[n++] = (a[w] * b[w]) / v[base + w];
    out[k] = orc_pairwise_sum(tmp, n);
  }
  free(tmp);
}

/* ---- np.correlate(x, m, 'same') for len(x)=N >= len(m)=W odd: stats/bayes.py:407,
 * general.pyx:210.  NumPy evaluates each lag with a BLAS dot whose internal order is
 * implementation defined; a plain ascending-index sum is used here. */
void orc_correlate_same(const double *x, long N, const double *m, int W, double *out)
{
  const int h = W / 2;
  for (long k = 0; k < N; k++) {
    long lo = k - h < 0 ? 0 : k - h, hi = k + h >= N ? N - 1 : k + h;
    double s = 0.;
    for (long i = lo; i <= hi; i++) s += x[i] * m[i - (k - h)];
    out[k] = s;
  }
}

/* ---- the sliding-window Bayes factors --------------------------------------------
 * stats/bayes.py:312-437 (tensors) + general.pyx:143-312 (per-sample assembly, adds
 * sumlogsigma) + log_marg_amp_full.c.  Priors / amplitude prior are added by the caller.
 *   d[N] raw data, v[N] noise variance, bg[P*W] background models, ms[G*W] templates,
 *   valid[G] (0 => reference fills the template with -inf => whole row is -inf),
 *   lnB[G*N] out.  If G == 0 only the background-only factor is produced (lnBbg).
 *   lnBbg[N] (may be NULL): bayes.py:439-563 + general.pyx:315-365.                   */
int orc_bayes_factors(const double *d, const double *v, long N, int W, int P,
                      const double *bg, const double *ms, const unsigned char *valid, int G,
                      int lhr, double *lnB, double *lnBbg)
{
  if (W % 2 == 0 || N < W || P + 1 > ORC_MAXNM) return -1;
  const int Nm = P + 1;
  double *dw = (double *)malloc(sizeof(double) * (size_t)N);
  double *bgcross = (double *)malloc(sizeof(double) * (size_t)(P * P) * (size_t)N);
  double *dbgr = (double *)malloc(sizeof(double) * (size_t)P * (size_t)N);
  double *lsk = (double *)malloc(sizeof(double) * (size_t)N);
  if (!dw || !bgcross || !dbgr || !lsk) return -2;

  for (int p = 0; p < P; p++)
    for (int q = p; q < P; q++)
      orc_window_cross(bg + p * W, bg + q * W, v, N, W, bgcross + ((size_t)(p * P + q)) * N);
  for (long i = 0; i < N; i++) dw[i] = d[i] / v[i];               /* bayes.py:403 */
  for (int p = 0; p < P; p++) orc_correlate_same(dw, N, bg + p * W, W, dbgr + (size_t)p * N);
  /* general.pyx:216-221: sumlogsigma = np.sum(np.log(sk[isfinite(sk)])), sk = sqrt(v) */
  long nfin = 0;
  for (long i = 0; i < N; i++) {
    double s = sqrt(v[i]);
    if (isfinite(s)) lsk[nfin++] = log(s);
  }
  const double sumlogsigma = orc_pairwise_sum(lsk, nfin);

  if (lnBbg) {
    for (long k = 0; k < N; k++) {
      double M[ORC_MAXNM * ORC_MAXNM], D[ORC_MAXNM];
      memset(M, 0, sizeof(M));
      for (int p = 0; p < P; p++) {
        D[p] = dbgr[(size_t)p * N + k];
        for (int q = p; q < P; q++) M[p * P + q] = bgcross[((size_t)(p * P + q)) * N + k];
      }
      lnBbg[k] = orc_marg_full(P, M, D, 1.0, 0) + sumlogsigma;
    }
  }

#pragma omp parallel for schedule(dynamic, 1)
  for (int g = 0; g < G; g++) {
    double *out = lnB + (size_t)g * N;
    if (!valid[g]) {
      for (long k = 0; k < N; k++) out[k] = -INFINITY;
      continue;
    }
    const double *m = ms + (size_t)g * W;
    double *mdcross = (double *)malloc(sizeof(double) * (size_t)N);
    double *mdbg = (double *)malloc(sizeof(double) * (size_t)P * (size_t)N);
    double *dm = (double *)malloc(sizeof(double) * (size_t)N);
    orc_window_cross(m, m, v, N, W, mdcross);
    for (int p = 0; p < P; p++) orc_window_cross(bg + p * W, m, v, N, W, mdbg + (size_t)p * N);
    orc_correlate_same(dw, N, m, W, dm);
    for (long k = 0; k < N; k++) {
      double M[ORC_MAXNM * ORC_MAXNM], D[ORC_MAXNM];
      memset(M, 0, sizeof(double) * (size_t)(Nm * Nm));
      for (int p = 0; p < P; p++) {
        for (int q = p; q < P; q++) M[p * Nm + q] = bgcross[((size_t)(p * P + q)) * N + k];
        M[p * Nm + P] = mdbg[(size_t)p * N + k];
        D[p] = dbgr[(size_t)p * N + k];
      }
      M[P * Nm + P] = mdcross[k];
      D[P] = dm[k];
      out[k] = orc_marg_full(Nm, M, D, 1.0, lhr) + sumlogsigma;
    }
    free(mdcross); free(mdbg); free(dm);
  }
  free(dw); free(bgcross); free(dbgr); free(lsk);
  return 0;
}

/* ---- parameter-estimation grid likelihood: stats/bayes.py:981-1055 +
 * general.pyx:610-656.  One chunk of L samples, un-whitened sums, scalar sigma.
 *   polyms[P*L], d[L], ms[G*L] -> post[G] (likelihood only; prior added by caller). */
int orc_pe_grid(const double *d, long L, int P, const double *polyms, const double *ms, long G,
                double sigma, double *post)
{
  const int Nm = P + 1;
  if (Nm > ORC_MAXNM) return -1;
  double *tmp0 = (double *)malloc(sizeof(double) * (size_t)L);
  double subdm[ORC_MAXNM], submm[ORC_MAXNM * ORC_MAXNM];
  memset(submm, 0, sizeof(submm));
  for (int p = 0; p < P; p++) {
    for (long i = 0; i < L; i++) tmp0[i] = polyms[p * L + i] * d[i];
    subdm[p] = orc_pairwise_sum(tmp0, L);
    for (int q = p; q < P; q++) {
      for (long i = 0; i < L; i++) tmp0[i] = polyms[p * L + i] * polyms[q * L + i];
      submm[p * Nm + q] = orc_pairwise_sum(tmp0, L);
    }
  }
  free(tmp0);
#pragma omp parallel
  {
    double *tmp = (double *)malloc(sizeof(double) * (size_t)L);
#pragma omp for schedule(static)
    for (long g = 0; g < G; g++) {
      const double *m = ms + (size_t)g * L;
      double M[ORC_MAXNM * ORC_MAXNM], D[ORC_MAXNM];
      memcpy(M, submm, sizeof(double) * (size_t)(Nm * Nm));
      memcpy(D, subdm, sizeof(double) * (size_t)P);
      for (int p = 0; p < P; p++) {
        for (long i = 0; i < L; i++) tmp[i] = m[i] * polyms[p * L + i];
        M[p * Nm + P] = orc_pairwise_sum(tmp, L);
      }
      for (long i = 0; i < L; i++) tmp[i] = m[i] * m[i];
      M[P * Nm + P] = orc_pairwise_sum(tmp, L);
      for (long i = 0; i < L; i++) tmp[i] = d[i] * m[i];
      D[P] = orc_pairwise_sum(tmp, L);
      post[g] = orc_marg_except_final(Nm, M, D, sigma);
    }
    free(tmp);
  }
  return 0;
}

/* ---- detector combine: finder/find.py:853-862 -------------------------------------
 * lnO[i] = Of[i] - logplus-fold(noise hypotheses), fold starts from -inf. */
void orc_combine(const double *Of, const double *noise, int nnoise, long N, long k0, long k1, double *lnO)
{
  for (long i = k0; i < k1; i++) {
    double denom = -INFINITY;
    for (int n = 0; n < nnoise; n++) denom = orc_logplus(denom, noise[(size_t)n * N + i]);
    lnO[i - k0] = nnoise > 0 ? Of[i] - denom : Of[i];
  }
}
