/* Test-infrastructure shim (oracle/): stands in for <gsl/gsl_math.h>, which the
 * reference includes (bayesflare/stats/log_marg_amp_full.h:7) only for M_LNPI.
 * GSL is not installed in this image. NOT part of the product path. */
#ifndef BFL_ORACLE_GSL_MATH_SHIM_H
#define BFL_ORACLE_GSL_MATH_SHIM_H
#include <math.h>
#ifndef M_LNPI
#define M_LNPI 1.14472988584940017414342735135 /* ln(pi) */
#endif
#endif
