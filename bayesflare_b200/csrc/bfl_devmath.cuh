// bfl_devmath.cuh -- device-side scalar helpers shared by the kernel translation units: phi(z) tables,
// e^d for d <= 0, the online log-sum-exp and logplus (general.pyx:519-547).  sm_100a only.
#pragma once
#include <math.h>

namespace bfl {

#define BFL_LN2 0.693147180559945309417232121458
#define BFL_LNPI 1.14472988584940017414342735135

__device__ __forceinline__ double neg_inf() { return __longlong_as_double(0xfff0000000000000ULL); }

// log(erfc(x)), finite for every finite x  (reference: gsl_sf_log_erfc, log_marg_amp_full.c:63)
__device__ __forceinline__ double log_erfc_dev(double x) {
  if (x <= 0.0) return log(erfc(x));
  return log(erfcx(x)) - x * x;
}

// phi(z) = z^2/2 + log erfc(-z/sqrt(2)):  the half-range amplitude integral in units of the
// amplitude signal-to-noise ratio z.
__device__ __forceinline__ double phi_half(double z) {
  const double t = -z * 0.70710678118654752440;
  if (t <= 0.0) return fma(0.5 * z, z, log(erfc(t)));
  return log(erfcx(t));
}

#include "phi_table.inc"

// phi(z) from the piecewise polynomial table (tools/gen_phi_table.py: degree 4 on intervals of 1/32,
// |error| < 1.3e-13).  `tab` is the shared-memory copy, one 16-byte aligned row of BFL_PHI_ROW doubles
// per interval, fetched with 128-bit loads (the look-up is lane-varying, so fewer, wider loads mean
// fewer shared-memory wavefronts).  Straight-line code: the interval index is clamped, and because
// the last interval is stored as the exact quadratic z^2/2 + log 2 (erfc = 2 to 1e-19 there) clamping
// extrapolates correctly for any larger z.  Only z < BFL_PHI_ZFAR needs the erfcx form
// (phi_half_far), which callers apply in a rarely-taken fix-up branch.  Full-range hypotheses need
// no table: their function is z^2/2.
#define BFL_PHI_STRIDE (BFL_PHI_ROW * BFL_PHI_NI)            // interval-major copy (epilogue kernel)
#define BFL_PHI_CM_SIZE ((BFL_PHI_DEG + 1) * BFL_PHI_NI)     // coefficient-major copy (single-kernel path)
#define BFL_PHI_ZFAR (BFL_PHI_ZMIN - 0.5 / BFL_PHI_INVH)
static_assert(BFL_PHI_DEG == 4 && BFL_PHI_ROW == 6, "phi_tab* are written for the degree-4 table");
// (a) coefficient-major shared-memory copy [coef][interval], 64-bit loads: best in the single-kernel path
__device__ __forceinline__ double phi_tab(double z, const double* __restrict__ tab) {
  const double s = fma(z, BFL_PHI_INVH, -(BFL_PHI_ZMIN) * BFL_PHI_INVH);
  int k = __double2int_rn(s);
  k = min(max(k, 0), BFL_PHI_NI - 1);
  const double x = s - (double)k;
  const double* c = tab + k;
  double v = c[4 * BFL_PHI_NI];
  v = fma(v, x, c[3 * BFL_PHI_NI]);
  v = fma(v, x, c[2 * BFL_PHI_NI]);
  v = fma(v, x, c[BFL_PHI_NI]);
  v = fma(v, x, c[0]);
  return v;
}
// (b) interval-major copy [interval][BFL_PHI_ROW] as generated, three 128-bit loads: fewer shared-memory
// wavefronts, best in the epilogue kernel (which is bound by these lane-varying look-ups)
__device__ __forceinline__ double phi_tab_rows(double z, const double* __restrict__ tab) {
  const double s = fma(z, BFL_PHI_INVH, -(BFL_PHI_ZMIN) * BFL_PHI_INVH);
  int k = __double2int_rn(s);
  k = min(max(k, 0), BFL_PHI_NI - 1);
  const double x = s - (double)k;
  const double2* c = reinterpret_cast<const double2*>(tab + k * BFL_PHI_ROW);
  const double2 c01 = c[0], c23 = c[1], c4 = c[2];
  double v = fma(c4.x, x, c23.y);
  v = fma(v, x, c23.x);
  v = fma(v, x, c01.y);
  v = fma(v, x, c01.x);
  return v;
}
static __device__ __noinline__ double phi_half_far(double z) {   // z < -12.06: log erfcx(-z/sqrt2)
  return log(erfcx(-z * 0.70710678118654752440));
}

#include "exphi_table.inc"
// E(z) = exp(phi(z)) for |z| <= BFL_EXPHI_ZC (tools/gen_exphi_table.py, relative error < 2e-12).
// The look-up is lane-varying and shared-memory wavefronts bound the epilogue, so a row is ONE 16-byte
// load (E_k, g_k = sqrt(2/pi)/E_k); the derivatives of phi at the node follow from dg/dz = -g dphi/dz and
// the rest is arithmetic: t = phi(z) - phi(z_k) to third order in d = z - z_k (|d| <= 1/256), E = E_k e^t
// with a degree-5 series (|t| < 0.025).  The caller checks |z| <= ZC; the index is clamped so an
// out-of-range z reads a valid row (its value is discarded).
#define BFL_EXPHI_SIZE (2 * BFL_EXPHI_NI)
__device__ __forceinline__ double exphi_eval(double z, const double* __restrict__ tab) {
  const double s = fma(z, BFL_EXPHI_INVH, BFL_EXPHI_ZC * BFL_EXPHI_INVH);
  int k = __double2int_rn(s);
  k = min(max(k, 0), BFL_EXPHI_NI - 1);
  const double d = (s - (double)k) * (1.0 / BFL_EXPHI_INVH);
  const double2 eg = reinterpret_cast<const double2*>(tab)[k];
  const double g = eg.y;
  const double p1 = (z - d) + g;                 // first derivative of phi at z_k
  const double p2 = fma(-g, p1, 1.0);            // second
  const double p3 = g * fma(p1, p1, -p2);        // third
  const double t = d * fma(d, fma(d, p3 * (1.0 / 6.0), 0.5 * p2), p1);
  double e = fma(t, 1.0 / 120.0, 1.0 / 24.0);
  e = fma(e, t, 1.0 / 6.0);
  e = fma(e, t, 0.5);
  e = fma(e, t, 1.0);
  e = fma(e, t, 1.0);
  return eg.x * e;
}

// e^d for d <= 0 (NaN propagates), two variants:
//  * exp_nonpos: d = (64 e + j) ln2/64 + r, |r| <= ln2/128; 2^(j/64) from a 64-entry shared-memory table
//    (`t64`), e^r from a degree-5 polynomial, 2^e patched into the exponent field (relative error
//    < 2e-15 where it matters).  Fewest FP64 instructions: used by the single-kernel path.
//  * exp_nonpos_poly: d = n ln2 + r, |r| <= ln2/2, degree-11 polynomial (relative error < 7e-15), no
//    table: used by the epilogue kernel, which is bound by lane-varying shared-memory look-ups.
__device__ __forceinline__ double exp_nonpos(double d, const double* __restrict__ t64) {
  d = (d < -700.0) ? -700.0 : d;
  const double MAGIC = 6755399441055744.0;   // 1.5 * 2^52
  const double t = fma(d, 92.332482616893656877, MAGIC);          // 64 / ln 2
  const int n = __double2loint(t);
  const double nf = t - MAGIC;
  const double r = fma(nf, -1.0830424696249145e-02, d);           // ln 2 / 64
  double p = 8.33333333333333333333e-03;                          // 1/5!
  p = fma(p, r, 4.16666666666666666667e-02);
  p = fma(p, r, 1.66666666666666666667e-01);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  p *= t64[n & 63];
  return __hiloint2double(__double2hiint(p) + ((n >> 6) << 20), __double2loint(p));
}

__device__ __forceinline__ double exp_nonpos_poly(double d) {
  d = (d < -700.0) ? -700.0 : d;
  const double MAGIC = 6755399441055744.0;
  const double t = fma(d, 1.4426950408889634074, MAGIC);
  const int n = __double2loint(t);
  const double nf = t - MAGIC;
  double r = fma(nf, -6.93147180369123816490e-01, d);
  r = fma(nf, -1.90821492927058770002e-10, r);
  double p = 2.50521083854417187751e-08;      // 1/11!
  p = fma(p, r, 2.75573192239858906526e-07);
  p = fma(p, r, 2.75573192239858906526e-06);
  p = fma(p, r, 2.48015873015873015873e-05);
  p = fma(p, r, 1.98412698412698412698e-04);
  p = fma(p, r, 1.38888888888888888889e-03);
  p = fma(p, r, 8.33333333333333333333e-03);
  p = fma(p, r, 4.16666666666666666667e-02);
  p = fma(p, r, 1.66666666666666666667e-01);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}

// online log-sum-exp, branch-free: m = running maximum (starts at LSE_EMPTY), s = sum exp(x_i - m)
#define LSE_EMPTY (-1.0e300)
__device__ __forceinline__ void lse_push(double& m, double& s, double x, const double* __restrict__ t64) {
  const double d = x - m;
  const double e = exp_nonpos(-fabs(d), t64);
  const bool up = __double2hiint(d) >= 0;     // sign bit (integer pipe); d == 0 gives the same sum either way
  s = fma(s, up ? e : 1.0, up ? 1.0 : e);
  m = up ? x : m;
}

__device__ __forceinline__ void lse_push_poly(double& m, double& s, double x) {
  const double d = x - m;
  const double e = exp_nonpos_poly(-fabs(d));
  const bool up = __double2hiint(d) >= 0;
  s = fma(s, up ? e : 1.0, up ? 1.0 : e);
  m = up ? x : m;
}

__device__ __forceinline__ double logplus_dev(double x, double y) {
  // general.pyx:519-547
  const bool bothinf = isinf(x) && isinf(y);
  if (bothinf && x < 0 && y < 0) return neg_inf();
  if (x > y && !bothinf) return x + log(1.0 + exp(y - x));
  if (x <= y && !bothinf) return y + log(1.0 + exp(x - y));
  return -neg_inf();
}

// ---- E(z) = exp(phi(z)) from the 8-byte table (tools/gen_exphi2_table.py), shared-memory access by 32-bit address ----
#include "exphi2_table.inc"
// shared-memory loads by 32-bit shared address (the table and the per-template records never change
// after the kernel's prologue, so the loads need not be volatile)
__device__ __forceinline__ double2 lds_f64x2(unsigned addr) {
  double2 v;
  asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ int4 lds_b32x4(unsigned addr) {
  int4 v;
  asm("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ int2 lds_b32x2(unsigned addr) {
  int2 v;
  asm("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ double lds_f64c(unsigned addr) {   // read-only data (the E table)
  double v;
  asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void cp_async_f64(unsigned saddr, const double* gaddr) {   // LDGSTS: global -> shared, 8 bytes
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" : : "r"(saddr), "l"(gaddr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" : : : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" : : : "memory"); }
__device__ __forceinline__ void sts_f64(unsigned addr, double v) {
  asm volatile("st.shared.f64 [%0], %1;" : : "r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ double lds_f64(unsigned addr) {   // volatile: the A-shape slice is rewritten per hypothesis
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}

__device__ __forceinline__ double rcp64_approx(double x) {   // one MUFU.RCP64H: ~20 good bits
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
}

// E(z) = E_k exp(t),  t = phi(z) - phi(z_k) to third order in d = z - z_k (|d| <= 1/256), e^t to fifth order.
// One 8-byte look-up; 26 FP64 instructions.
__device__ __forceinline__ double exphi2_one(double z6, unsigned tab, unsigned& kmax) {
  const double MAGIC = 6755399441055744.0;   // 1.5 * 2^52: the low word of x + MAGIC is rint(x)
  const double s = fma(z6, BFL_EXPHI2_INVH, MAGIC);
  const unsigned k = (unsigned)__double2loint(s);
  kmax = max(kmax, k);
  const double Ek = lds_f64c(tab + 8u * min(k, (unsigned)(BFL_EXPHI2_NI - 1)));
  const double kf = s - MAGIC;
  const double d = fma(kf, -1.0 / BFL_EXPHI2_INVH, z6);
  const double r0 = rcp64_approx(Ek);
  const double r = fma(r0, fma(-Ek, r0, 1.0), r0);                    // 1 / E_k
  const double g = r * 0.79788456080286535588;                         // sqrt(2/pi) / E_k
  const double p1 = fma(kf, 1.0 / BFL_EXPHI2_INVH, -BFL_EXPHI2_ZC) + g; // first derivative of phi at z_k
  const double p2 = fma(-g, p1, 1.0);                                  // second
  const double p3 = g * fma(p1, p1, -p2);                              // third
  const double t = d * fma(d, fma(d, p3 * (1.0 / 6.0), 0.5 * p2), p1);
  double e = fma(t, 1.0 / 120.0, 1.0 / 24.0);
  e = fma(e, t, 1.0 / 6.0);
  e = fma(e, t, 0.5);
  e = fma(e, t, 1.0);
  e = fma(e, t, 1.0);
  return Ek * e;
}

}  // namespace bfl
