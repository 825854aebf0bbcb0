/* Test-infrastructure shim (oracle/): stands in for <gsl/gsl_sf_erf.h>, which the
 * reference uses for gsl_sf_log_erfc (bayesflare/stats/log_marg_amp_full.c:63,
 * general.pyx:462) and gsl_sf_erf (general.pyx:137). GSL is a third-party,
 * un-vendored, un-pinned dependency of the reference (setup.py:42-43 shells out
 * to gsl-config) and is absent from this image, so parity at this boundary is
 * anchored on the mathematical definition log(erfc(x)) evaluated to ~1 ulp:
 *   x <= 25 : log(erfc(x)) with glibc's erfc (result is a normal double there)
 *   x  > 25 : -x^2 - log(x sqrt(pi)) + log(asymptotic series), |term ratio| < 1e-3
 * NOT part of the product path. */
#ifndef BFL_ORACLE_GSL_SF_ERF_SHIM_H
#define BFL_ORACLE_GSL_SF_ERF_SHIM_H
#include <math.h>

static inline double gsl_sf_erf(double x) { return erf(x); }

static inline double gsl_sf_log_erfc(double x)
{
  if (!(x > 25.0)) { return log(erfc(x)); }
  /* erfc(x) ~ exp(-x^2)/(x sqrt(pi)) * sum_n (-1)^n (2n-1)!! / (2x^2)^n */
  {
    const double r = 1.0 / (2.0 * x * x);
    double term = 1.0, sum = 1.0;
    int n;
    for (n = 1; n < 30; n++) {
      term *= -(2.0 * n - 1.0) * r;
      sum += term;
      if (fabs(term) < 1e-18 * fabs(sum)) break;
    }
    return -x * x - log(x) - 0.5 * 1.14472988584940017414342735135 + log(sum);
  }
}
#endif
