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

// e = exp(-tau), ome = 1 - exp(-tau) for tau >= 0.  Relative error of e ~2e-16; ome is free of cancellation.
__device__ __forceinline__ void exp_neg(double tau, double& e, double& ome) {
  const double kMagic = 6755399441055744.0;  // 1.5 * 2^52: adding it rounds to nearest integer
  double t = fma(tau, -1.4426950408889634074, kMagic);
  int n = __double2loint(t);
  double fn = t - kMagic;
  double r = fma(fn, -6.93147180369123816490e-01, -tau);
  r = fma(fn, -1.90821492927058770002e-10, r);
  // exp(r) - 1 - r = r^2 * q(r), |r| <= ln2/2; Taylor through r^12
  double q = 2.08767569878680989792e-09;          // 1/12!
  q = fma(q, r, 2.50521083854417187751e-08);      // 1/11!
  q = fma(q, r, 2.75573192239858906526e-07);      // 1/10!
  q = fma(q, r, 2.75573192239858906526e-06);      // 1/9!
  q = fma(q, r, 2.48015873015873015873e-05);      // 1/8!
  q = fma(q, r, 1.98412698412698412698e-04);      // 1/7!
  q = fma(q, r, 1.38888888888888888889e-03);      // 1/6!
  q = fma(q, r, 8.33333333333333333333e-03);      // 1/5!
  q = fma(q, r, 4.16666666666666666667e-02);      // 1/4!
  q = fma(q, r, 1.66666666666666666667e-01);      // 1/3!
  q = fma(q, r, 0.5);
  double em1 = fma(r * r, q, r);                  // exp(r) - 1
  double s = 1.0 + em1;                           // in [0.70, 1.42]
  e = __hiloint2double(__double2hiint(s) + (n << 20), __double2loint(s));  // s * 2^n, valid while normal
  ome = (n == 0) ? -em1 : 1.0 - e;
  if (!(tau < 700.0)) {                           // subnormal / zero / NaN results: rare, take the library path
    e = exp(-tau);
    ome = 1.0 - e;
  }
}

struct SegResult {
  double Iout, J;
};

// kpos: kappa > 0.  invtau = 1 / (kappa * dpath) (only used when kpos).
template <bool FAITHFUL>
__device__ __forceinline__ SegResult segment_update(double Iin, double kappa, double dpath, double invtau, bool kpos,
                                                    double nseg_unused = 0.) {
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
    double phi = kpos ? ome * invtau : 1.0;
    // Iout == 0 (underflow): the reference gets (Iin - 0)/log(inf) = 0
    long long bits = __double_as_longlong(r.Iout);
    r.J = ((bits << 1) == 0) ? 0.0 : Iin * phi;
  }
  return r;
}

}  // namespace rtb
