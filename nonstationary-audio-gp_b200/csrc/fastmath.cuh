// Branch-free FP64 elementary functions for the latency-critical sequential passes.
//
// The CUDA library versions of 1/x, sqrt, rsqrt, exp and log are correctly handled
// for every special case, which costs slow-path branches that ptxas will not
// interleave across independent evaluations.  On the ADF critical path the
// arguments are known to be ordinary numbers, so these versions start from the
// hardware seed (MUFU.RCP64H / MUFU.RSQ64H, ~20 bits) and refine with straight-line
// FMAs.  Accuracy is <= 2 ulp over the stated domains (tests/test_gpu_mom.py checks
// them against numpy through nsagp_fastmath_eval); the parity budget of the path is
// 1e-8 relative.
#pragma once
#include "common.cuh"

namespace nsagp {

// 1/x for finite normal x (either sign).  Not for 0, Inf, NaN, subnormals.
__device__ __forceinline__ double rcp_fast(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

// 1/sqrt(x) for finite normal x > 0.
__device__ __forceinline__ double rsqrt_fast(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double hx = 0.5 * x;
  double e = fma(-hx * r, r, 0.5);         // (1 - x r^2) / 2
  r = fma(r, e, r);
  e = fma(-hx * r, r, 0.5);
  r = fma(r, e, r);
  e = fma(-hx * r, r, 0.5);
  return fma(r, e, r);
}

// sqrt(x) for x >= 0 finite (returns +0 for 0; NaN for negative x, as sqrt does).
__device__ __forceinline__ double sqrt_fast(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double g = x * r;                         // ~ sqrt(x)
  double h = 0.5 * r;                       // ~ 1 / (2 sqrt(x))
  double e = fma(-h, g, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  e = fma(-h, g, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  const double d = fma(-g, g, x);           // residual
  g = fma(d, h, g);
  return (x == 0.0) ? 0.0 : g;              // 0 * Inf = NaN above; negative x stays NaN
}

// Variants with one Newton step fewer (nsagp_fastmath_eval ops 6-8 measure them).
__device__ __forceinline__ double rcp_fast2(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}
__device__ __forceinline__ double rsqrt_fast2(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double hx = 0.5 * x;
  double e = fma(-hx * r, r, 0.5);
  r = fma(r, e, r);
  e = fma(-hx * r, r, 0.5);
  return fma(r, e, r);
}
__device__ __forceinline__ double sqrt_fast2(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double g = x * r;
  double h = 0.5 * r;
  const double e = fma(-h, g, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  const double d = fma(-g, g, x);
  g = fma(d, h, g);
  return (x == 0.0) ? 0.0 : g;
}

// exp(x), straight-line, for finite x.  The argument is clamped to [-708, 709]: below, the
// result is exp(-708) = 3.3e-308 instead of a subnormal or 0; above, exp(709) = 8.2e307 instead
// of Inf.  CHECKED = true additionally returns exactly 0 below -708 and propagates NaN.
template <bool CHECKED>
__device__ __forceinline__ double exp_fast_t(double x) {
  const double xc = fmin(fmax(x, -708.0), 709.0);
  const double t = fma(xc, 1.4426950408889634074, 6755399441055744.0);      // log2(e); 1.5 * 2^52 rounds to nearest integer
  const int n = __double2loint(t);
  const double fn = t - 6755399441055744.0;
  double r = fma(fn, -6.93147180369123816490e-01, xc);
  r = fma(fn, -1.90821492927058770002e-10, r);
  // exp(r), |r| <= ln2/2 : Taylor to degree 13 (truncation 4e-18), Estrin evaluation
  const double r2 = r * r;
  const double r4 = r2 * r2;
  const double r8 = r4 * r4;
  const double p01 = 1.0 + r;
  const double p23 = fma(r, (1.0 / 6.0), 0.5);
  const double p45 = fma(r, (1.0 / 120.0), (1.0 / 24.0));
  const double p67 = fma(r, (1.0 / 5040.0), (1.0 / 720.0));
  const double p89 = fma(r, (1.0 / 362880.0), (1.0 / 40320.0));
  const double pab = fma(r, (1.0 / 39916800.0), (1.0 / 3628800.0));
  const double pcd = fma(r, (1.0 / 6227020800.0), (1.0 / 479001600.0));
  const double q0 = fma(r2, p23, p01);
  const double q1 = fma(r2, p67, p45);
  const double q2 = fma(r2, pab, p89);
  const double s0 = fma(r4, q1, q0);
  const double s1 = fma(r4, pcd, q2);
  const double p = fma(r8, s1, s0);
  const double res = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
  if (!CHECKED) return res;
  return (x <= -708.0) ? 0.0 : ((x != x) ? x : res);
}
__device__ __forceinline__ double exp_fast(double x) { return exp_fast_t<true>(x); }

// The same for finite x with the range handling OFF the dependent chain: the polynomial runs on the unclamped argument
// (for x outside [-708, 709] the exponent arithmetic wraps and the value is garbage) and two selects, whose predicates
// depend on x alone, put the saturated values in at the end.  Two instructions shorter at the head of the chain.
__device__ __forceinline__ double exp_fast_sel(double x) {
  const double t = fma(x, 1.4426950408889634074, 6755399441055744.0);
  const int n = __double2loint(t);
  const double fn = t - 6755399441055744.0;
  double r = fma(fn, -6.93147180369123816490e-01, x);
  r = fma(fn, -1.90821492927058770002e-10, r);
  const double r2 = r * r;
  const double r4 = r2 * r2;
  const double r8 = r4 * r4;
  const double p01 = 1.0 + r;
  const double p23 = fma(r, (1.0 / 6.0), 0.5);
  const double p45 = fma(r, (1.0 / 120.0), (1.0 / 24.0));
  const double p67 = fma(r, (1.0 / 5040.0), (1.0 / 720.0));
  const double p89 = fma(r, (1.0 / 362880.0), (1.0 / 40320.0));
  const double pab = fma(r, (1.0 / 39916800.0), (1.0 / 3628800.0));
  const double pcd = fma(r, (1.0 / 6227020800.0), (1.0 / 479001600.0));
  const double q0 = fma(r2, p23, p01);
  const double q1 = fma(r2, p67, p45);
  const double q2 = fma(r2, pab, p89);
  const double s0 = fma(r4, q1, q0);
  const double s1 = fma(r4, pcd, q2);
  const double p = fma(r8, s1, s0);
  double res = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
  res = (x < -708.0) ? 3.3075530140267002e-308 : res;          // exp(-708), as exp_fast_t<false>
  res = (x > 709.0) ? 8.2184074615549724e+307 : res;           // exp(709)
  return res;
}

// log(u) for finite normal u >= 1 (the softplus link's 1 + exp(.)); fdlibm's scheme:
// u = 2^k (1+f), s = f/(2+f), log(1+f) = f - f^2/2 + s (f^2/2 + R(s^2)).
// CHECKED = true propagates NaN.
template <bool CHECKED>
__device__ __forceinline__ double log_ge1_fast_t(double u) {
  int hi = __double2hiint(u);
  const int lo = __double2loint(u);
  int k = (hi >> 20) - 1023;
  hi &= 0x000fffff;
  const int i = (hi + 0x95f64) & 0x100000;  // mantissa above sqrt(2): halve it, bump the exponent
  k += i >> 20;
  const double m = __hiloint2double(hi | (i ^ 0x3ff00000), lo);
  const double f = m - 1.0;
  const double s = f * rcp_fast2(2.0 + f);
  const double z = s * s;
  const double w = z * z;
  const double t1 = w * fma(w, fma(w, 1.531383769920937332e-01, 2.222219843214978396e-01), 3.999999999940941908e-01);
  const double t2 = z * fma(w, fma(w, fma(w, 1.479819860511658591e-01, 1.818357216161805012e-01), 2.857142874366239149e-01), 6.666666666666735130e-01);
  const double R = t2 + t1;
  const double hfsq = 0.5 * f * f;
  const double dk = (double)k;
  const double res = fma(dk, 6.93147180369123816490e-01, f - (hfsq - fma(s, hfsq + R, dk * 1.90821492927058770002e-10)));
  if (!CHECKED) return res;
  return (u != u) ? u : res;
}
__device__ __forceinline__ double log_ge1_fast(double u) { return log_ge1_fast_t<true>(u); }

// log(u) for finite normal u >= 1 WITHOUT the division of the scheme above: u = 2^k m, the top seven mantissa bits
// j select c_j = 1 + j/128, r = m * fl(1/c_j) - 1 in [0, 2^-7) (one FMA; the rounding of 1/c_j is folded into the
// tabulated -log(fl(1/c_j))), log(m) = log1p(r) - log(fl(1/c_j)) with a degree-8 Taylor polynomial (truncation r^9/9 <=
// 2^-66).  c_0 = 1 and its table entry is exactly 0, so results near 0 keep their relative accuracy.  Dependent chain:
// ~100 cycles against ~200 (the reciprocal alone is ~70); <= 1.3 ulp (tests/test_gpu_mom.py).  The table (2 KB) must
// be in shared memory (log_tab_fill) -- the lanes of a warp read different rows.
__device__ const double kLogTab[256] = {
#include "logtab.inc"
};
constexpr int kLogTabDoubles = 256;

__device__ __forceinline__ void log_tab_fill(double* s_tab, int tid, int nthreads) {
  for (int i = tid; i < kLogTabDoubles; i += nthreads) s_tab[i] = kLogTab[i];
}

// tab_saddr: shared-memory BYTE address of the table (16-byte aligned)
__device__ __forceinline__ double log_ge1_tab(double u, unsigned tab_saddr) {
  const int hi = __double2hiint(u);
  const int lo = __double2loint(u);
  const int k = (hi >> 20) - 1023;
  const unsigned j = ((unsigned)hi >> 13) & 127u;
  double ic, lc;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(ic), "=d"(lc) : "r"(tab_saddr + 16u * j));
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
  const double dk = (double)k;
  const double r = fma(m, ic, -1.0);
  const double A = fma(dk, 6.93147180369123816490e-01, lc);
  const double r2 = r * r;
  const double b1 = fma(r, (1.0 / 3.0), -0.5);
  const double b2 = fma(r, 0.2, -0.25);
  const double b3 = fma(r, (1.0 / 7.0), -(1.0 / 6.0));
  const double r4 = r2 * r2;
  const double q01 = fma(r2, b2, b1);
  const double q23 = fma(r2, -0.125, b3);
  const double q = fma(r4, q23, q01);
  const double B = fma(dk, 1.90821492927058770002e-10, fma(r2, q, r));
  return A + B;
}

// The reference's link, literally log(1 + exp(g - shift)) (likModulatorNMFPower.m:44 with
// link = @(g) log(1+exp(g-c))): the rounding of 1 + exp(.) is part of the arithmetic.
__device__ __forceinline__ double softplus_fast(double xs) {
  return log_ge1_fast(1.0 + exp_fast(xs));
}
// For finite arguments only (no NaN propagation): the moment warps of the sequential passes.
__device__ __forceinline__ double softplus_fast_finite(double xs) {
  return log_ge1_fast_t<false>(1.0 + exp_fast_t<false>(xs));
}

// The same with the table-driven logarithm (the moment warps of the CTA kernels: the link is the head of their chain).
__device__ __forceinline__ double softplus_tab_finite(double xs, unsigned tab_saddr) {
  return log_ge1_tab(1.0 + exp_fast_sel(xs), tab_saddr);
}

// sqrt(x) for finite normal x > 0 (no zero handling).
__device__ __forceinline__ double sqrt_fast2_pos(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double g = x * r;
  double h = 0.5 * r;
  const double e = fma(-h, g, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  const double d = fma(-g, g, x);
  return fma(d, h, g);
}

}  // namespace nsagp
