// Per-segment arithmetic of the diffuse sweep (fp64, no tensor cores: the update is a chain of scalar
// exponentials, not a contraction).
//
// Reference (transportRoutinesModule.f90:651-698 and :1036-1054, identical inline copy equiSources.f90:1611-1643):
//     tau  = kappa * dpath;  Iout = Iin * exp(-tau)
//     Jseg = (Iin - Iout) / log(Iin / Iout)   if Iout < Iin,   else 0.5 * (Iin + Iout)
// FAITHFUL mode evaluates exactly that sequence (CUDA libm exp/log, IEEE division).
// FAST mode uses log(Iin/Iout) == tau, i.e. Jseg = Iin * (1 - exp(-tau)) / tau, with (1 - exp(-tau)) obtained
// without cancellation from the same polynomial that yields exp(-tau): one exponential and no log / division per
// segment.  The two differ by the rounding noise of the reference formula, ~1.1e-16 / tau relative.
#pragma once
#include <cuda_runtime.h>

namespace rtb {

// Coefficients live in constant memory so that every DFMA takes its coefficient straight from the constant bank
// (no register moves for 64-bit immediates; the kernel is issue-slot bound otherwise).
static __constant__ double kExpC[16] = {
    2.08767569878680989792e-09,  // 1/12!
    2.50521083854417187751e-08,  // 1/11!
    2.75573192239858906526e-07,  // 1/10!
    2.75573192239858906526e-06,  // 1/9!
    2.48015873015873015873e-05,  // 1/8!
    1.98412698412698412698e-04,  // 1/7!
    1.38888888888888888889e-03,  // 1/6!
    8.33333333333333333333e-03,  // 1/5!
    4.16666666666666666667e-02,  // 1/4!
    1.66666666666666666667e-01,  // 1/3!
    0.5,
    -1.4426950408889634074,      // [11] -log2(e)
    6755399441055744.0,          // [12] 1.5 * 2^52: adding it rounds to nearest integer
    -6.93147180369123816490e-01, // [13] -ln2 hi
    -1.90821492927058770002e-10, // [14] -ln2 lo
    1400.0};                     // [15] clamp: 2^(n/2) must stay a normal number

// e = exp(-tau), ome = 1 - exp(-tau) for tau >= 0.  Relative error of e ~2e-16; ome is free of cancellation.
// Straight-line code on purpose: a rare-case branch here would end the basic block and keep the compiler from
// interleaving the (independent) exponentials of the three frequency groups and of the 1..3 segments, which is where
// the instruction-level parallelism of the sweep comes from.
//   2^n is applied as two factors so that results down to the subnormal range and an underflow to 0 (tau > 745)
//   come out of the same multiplications; 1 - sc*(1 + em1) = fma(-sc, em1, 1 - sc) is exactly -em1 for n == 0.
__device__ __forceinline__ void exp_neg(double tau, double& e, double& ome) {
  tau = fmin(tau, kExpC[15]);                                     // beyond this exp(-tau) is 0 in fp64 anyway
  double t = fma(tau, kExpC[11], kExpC[12]);
  int n = __double2loint(t);
  double fn = t - kExpC[12];
  double r = fma(fn, kExpC[13], -tau);
  r = fma(fn, kExpC[14], r);
  // exp(r) - 1 - r = r^2 * q(r), |r| <= ln2/2; Taylor through r^12
  double q = kExpC[0];
#pragma unroll
  for (int i = 1; i <= 10; i++) q = fma(q, r, kExpC[i]);
  double em1 = fma(r * r, q, r);                                  // exp(r) - 1
  int h = n >> 1;
  double sc1 = __hiloint2double(0x3ff00000 + (h << 20), 0);       // 2^h
  double sc2 = __hiloint2double(0x3ff00000 + ((n - h) << 20), 0); // 2^(n-h)
  double sc = sc1 * sc2;                                          // 2^n (0 or subnormal below 2^-1022)
  ome = fma(-sc, em1, 1.0 - sc);
  e = ((1.0 + em1) * sc1) * sc2;
}

// exp(-tau) only (upstream segments that are recomputed need no J)
__device__ __forceinline__ double exp_neg_only(double tau) {
  tau = fmin(tau, kExpC[15]);
  double t = fma(tau, kExpC[11], kExpC[12]);
  int n = __double2loint(t);
  double fn = t - kExpC[12];
  double r = fma(fn, kExpC[13], -tau);
  r = fma(fn, kExpC[14], r);
  double q = kExpC[0];
#pragma unroll
  for (int i = 1; i <= 10; i++) q = fma(q, r, kExpC[i]);
  double s = 1.0 + fma(r * r, q, r);
  int h = n >> 1;
  double sc1 = __hiloint2double(0x3ff00000 + (h << 20), 0);
  double sc2 = __hiloint2double(0x3ff00000 + ((n - h) << 20), 0);
  return (s * sc1) * sc2;
}

struct SegResult {
  double Iout, J;
};

// FAST mode expects kappa > 0 (callers replace an exact zero by a tiny positive number, for which every formula
// below returns the kappa = 0 limits Iout = Iin, J = Iin) and invtau = 1 / (kappa * dpath).
template <bool FAITHFUL>
__device__ __forceinline__ SegResult segment_update(double Iin, double kappa, double dpath, double invtau) {
  SegResult r;
  if (FAITHFUL) {
    double tau = __dmul_rn(kappa, dpath);
    double a = exp(-tau);
    r.Iout = __dmul_rn(Iin, a);
    if (r.Iout < Iin) r.J = __ddiv_rn(__dsub_rn(Iin, r.Iout), log(__ddiv_rn(Iin, r.Iout)));
    else r.J = __dmul_rn(0.5, __dadd_rn(Iin, r.Iout));
  } else {
    double tau = kappa * dpath, e, ome;
    exp_neg(tau, e, ome);
    r.Iout = Iin * e;
    // Iout == 0 (underflow): the reference gets (Iin - 0)/log(Iin/0) = 0.  (Where Iout is a non-zero SUBNORMAL,
    // i.e. per-segment tau of ~650-745, the reference's log sees only the few bits Iout has left; FAST mode returns
    // the smooth value there -- use FAITHFUL mode to reproduce that artefact.)
    r.J = ((__double_as_longlong(r.Iout) << 1) == 0) ? 0.0 : Iin * (ome * invtau);
  }
  return r;
}

}  // namespace rtb
