// Portable fp64 exp / log built only from IEEE-754 +, *, /, fma and integer bit operations, so that the SAME
// source gives bit-identical results in a CUDA kernel (sm_100a) and in host code compiled with a hardware or
// correctly-rounded software fma (std::fma is correctly rounded by the C standard).
//
// Why it exists: the point-source deposits of the reference are differences R(d) - R(d + tau) of exponentials of
// interpolated logarithms (equiSources.f90:3247-3260, 4205-4238).  For a short segment the difference cancels, and a
// last-bit disagreement between two libm implementations (glibc on the host, CUDA libm on the device) is amplified
// by ~1/tau: the comparison "GPU vs CPU restatement" would then measure libm noise instead of the algorithm.  With
// these functions on both sides (RTB200_MATH_FAITHFUL on the device, the `portable` switch of the CPU oracle) every
// deposit is bit-identical and only the summation order of the per-cell accumulation differs.
// Accuracy: < 1 ulp (exp), < 1.5 ulp (log) on the ranges used here; special cases: exp(x < -745) = 0,
// exp(x > 709.78) = +inf, log(0) = -inf, log(x < 0) = NaN, subnormal arguments of log are handled.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define RTB_HD __host__ __device__ __forceinline__
#else
#define RTB_HD inline
#endif

namespace rtb_pm {

RTB_HD double pm_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}
RTB_HD double pm_mul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  return a * b;  // host translation units are compiled with -ffp-contract=off
#endif
}
RTB_HD double pm_add(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
RTB_HD double pm_div(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __ddiv_rn(a, b);
#else
  return a / b;
#endif
}
RTB_HD int64_t pm_bits(double x) {
#if defined(__CUDA_ARCH__)
  return __double_as_longlong(x);
#else
  int64_t i; std::memcpy(&i, &x, 8); return i;
#endif
}
RTB_HD double pm_from_bits(int64_t i) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double(i);
#else
  double x; std::memcpy(&x, &i, 8); return x;
#endif
}

// exp(x) = 2^k * exp(r), k = nearest integer to x / ln2, r = x - k ln2 (two-part ln2), |r| <= 0.3466;
// exp(r) by Horner through r^13 / 13!  (truncation 0.3466^14 / 14! = 4e-18).
RTB_HD double pm_exp(double x) {
  if (x != x) return x;
  if (x > 709.782712893384) return pm_from_bits(0x7ff0000000000000LL);
  if (x < -745.2) return 0.0;
  const double kShift = 6755399441055744.0;  // 1.5 * 2^52
  const double t = pm_fma(x, 1.4426950408889634074, kShift);
  const double fk = pm_add(t, -kShift);
  const int kk = (int)(int32_t)(uint32_t)(pm_bits(t) & 0xffffffffLL);  // low word of t holds k (two's complement)
  double r = pm_fma(fk, -6.93147180369123816490e-01, x);
  r = pm_fma(fk, -1.90821492927058770002e-10, r);
  double q = 1.6059043836821614599e-10;  // 1/13!
  q = pm_fma(q, r, 2.0876756987868098979e-09);
  q = pm_fma(q, r, 2.5052108385441718775e-08);
  q = pm_fma(q, r, 2.7557319223985890653e-07);
  q = pm_fma(q, r, 2.7557319223985890653e-06);
  q = pm_fma(q, r, 2.4801587301587301587e-05);
  q = pm_fma(q, r, 1.9841269841269841270e-04);
  q = pm_fma(q, r, 1.3888888888888888889e-03);
  q = pm_fma(q, r, 8.3333333333333333333e-03);
  q = pm_fma(q, r, 4.1666666666666666667e-02);
  q = pm_fma(q, r, 1.6666666666666666667e-01);
  q = pm_fma(q, r, 0.5);
  const double r2 = pm_mul(r, r);
  const double em1 = pm_fma(r2, q, r);          // exp(r) - 1
  // scale by 2^kk in two steps so that results in the subnormal range round once, from a normal number
  const int k1 = kk / 2, k2 = kk - k1;
  const double s1 = pm_from_bits((int64_t)(1023 + k1) << 52), s2 = pm_from_bits((int64_t)(1023 + k2) << 52);
  const double y = pm_fma(s1, em1, s1);         // 2^k1 * exp(r)
  return pm_mul(y, s2);
}

// log(x): x = 2^e * m with m in [sqrt(1/2), sqrt(2)); s = (m - 1) / (m + 1), log m = 2 atanh(s) =
// 2 s (1 + s^2/3 + s^4/5 + ... ), |s| <= 0.1716, series through s^24 (truncation 2e-20); result e*ln2_hi + (log m +
// e*ln2_lo) with a two-part ln2 whose high part has 32 trailing zero bits (e * ln2_hi is exact).
RTB_HD double pm_log(double x) {
  if (x != x) return x;
  if (x < 0.0) return pm_from_bits(0x7ff8000000000000LL);
  if (x == 0.0) return pm_from_bits((int64_t)0xfff0000000000000ULL);
  int64_t b = pm_bits(x);
  if (b == 0x7ff0000000000000LL) return x;
  int e = 0;
  if (b < 0x0010000000000000LL) {  // subnormal: scale by 2^54
    x = pm_mul(x, 18014398509481984.0);
    b = pm_bits(x);
    e = -54;
  }
  e += (int)(b >> 52) - 1023;
  int64_t mb = (b & 0x000fffffffffffffLL) | 0x3ff0000000000000LL;  // m in [1, 2)
  if (mb >= 0x3ff6a09e667f3bcdLL) {  // m >= sqrt(2): halve
    mb -= 0x0010000000000000LL;
    e += 1;
  }
  const double m = pm_from_bits(mb);
  const double num = pm_add(m, -1.0);  // exact
  const double den = pm_add(m, 1.0);
  const double s = pm_div(num, den);
  const double z = pm_mul(s, s);
  double p = 8.0e-02;                  // 2/25
  p = pm_fma(p, z, 8.6956521739130432e-02);  // 2/23
  p = pm_fma(p, z, 9.5238095238095233e-02);  // 2/21
  p = pm_fma(p, z, 1.0526315789473684e-01);  // 2/19
  p = pm_fma(p, z, 1.1764705882352941e-01);  // 2/17
  p = pm_fma(p, z, 1.3333333333333333e-01);  // 2/15
  p = pm_fma(p, z, 1.5384615384615385e-01);  // 2/13
  p = pm_fma(p, z, 1.8181818181818182e-01);  // 2/11
  p = pm_fma(p, z, 2.2222222222222221e-01);  // 2/9
  p = pm_fma(p, z, 2.8571428571428570e-01);  // 2/7
  p = pm_fma(p, z, 4.0000000000000002e-01);  // 2/5
  p = pm_fma(p, z, 6.6666666666666663e-01);  // 2/3
  // log m = 2 s + s * z * p.  2 s carries the rounding error of the division; recover it:
  // s = num/den - eps, eps from the fma residual  num - s*den
  const double resid = pm_fma(-s, den, num);         // exact remainder of the division
  const double corr = pm_div(resid, den);            // s_true = s + corr
  const double hi = pm_mul(2.0, s);
  double lo = pm_fma(pm_mul(s, z), p, pm_mul(2.0, corr));
  const double fe = (double)e;
  lo = pm_fma(fe, 1.90821492927058770002e-10, lo);
  return pm_add(pm_mul(fe, 6.93147180369123816490e-01), pm_add(hi, lo));
}

}  // namespace rtb_pm
