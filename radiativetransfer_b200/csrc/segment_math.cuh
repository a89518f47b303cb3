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

// exp(-tau) = 2^(-k/16) * exp(-r),  k = round(tau * 16/ln2),  r = tau - k*ln2/16,  |r| <= ln2/32:
// a 16-entry table of 2^(-j/16) (one row of shared memory: any set of lanes reads it without bank conflicts), an
// exponent-field subtraction for 2^-(k>>4) and a degree-7 polynomial.  13-15 FP64 instructions instead of ~20 for a
// table-free evaluation; the sweep is bound by the FP64 pipe, so this is where its time goes.
// Constants sit in constant memory so that the DFMAs read them from the constant bank / uniform registers.
static __constant__ double kExpC[16] = {
    -1.98412698412698412698e-04,  // [0] -1/7!
    1.38888888888888888889e-03,   // [1]  1/6!
    -8.33333333333333333333e-03,  // [2] -1/5!
    4.16666666666666666667e-02,   // [3]  1/4!
    -1.66666666666666666667e-01,  // [4] -1/3!
    0.5,                          // [5]
    -1.0,                         // [6]
    23.083120654223414,           // [7]  16/ln2
    6755399441055744.0,           // [8]  1.5 * 2^52: adding it rounds to nearest integer
    -0.0433216979727149,          // [9]  -ln2/16, high part (27 trailing zero bits: k*hi is exact)
    -8.12281680868118e-10,        // [10] -ln2/16, low part
    707.0,                        // [11] clamp: keeps 2^-(k>>4) * 2^(-j/16) a normal number
    0., 0., 0., 0.};
static __constant__ double kExpTable[16] = {
    1.00000000000000000000e+00, 9.57603280698573700036e-01, 9.17004043204671215328e-01, 8.78126080186649726755e-01,
    8.40896415253714502036e-01, 8.05245165974627141736e-01, 7.71105412703970372057e-01, 7.38413072969749673113e-01,
    7.07106781186547572737e-01, 6.77127773468446325644e-01, 6.48419777325504820276e-01, 6.20928906036742001007e-01,
    5.94603557501360513449e-01, 5.69394317378345782288e-01, 5.45253866332628844837e-01, 5.22136891213706877402e-01};

// Straight-line code on purpose: a rare-case branch here would end the basic block and keep the compiler from
// interleaving the (independent) exponentials of the three frequency groups and of the 1..3 segments, which is where
// the instruction-level parallelism of the sweep comes from.  tau is clamped at 707: exp(-tau) below 2^-1020 is
// returned as ~2^-1020 (it only ever multiplies an intensity, and such products are far below anything physical).
// Returns Ts = 2^(-k/16) and p = exp(-r) - 1, so that  exp(-tau) = Ts + Ts*p  and  1 - exp(-tau) = (1 - Ts) - Ts*p
// (1 - Ts is exact, and for k == 0 the latter is exactly -p: no cancellation for small tau).
template <bool GUARD = true>
__device__ __forceinline__ void exp_neg_parts(double tau, const double* __restrict__ T, double& Ts, double& p) {
  // tau >= 0 (kappa >= 0, path > 0): clamp at 707 with ONE integer min on the high word instead of an FP64 compare
  // and two selects (0x40861800'00000000 = 707.0; the low word stays, so the clamped value lies in [707, 707.0005)).
  // GUARD = false: the caller knows tau <= 64 for the whole warp (see sweep_cell_kernel) and skips the clamp.
  if (GUARD) tau = __hiloint2double(min(__double2hiint(tau), 0x40861800), __double2loint(tau));
  double t = fma(tau, kExpC[7], kExpC[8]);
  int k = __double2loint(t);
  double fn = t - kExpC[8];
  double r = fma(fn, kExpC[9], tau);
  r = fma(fn, kExpC[10], r);
  double q = kExpC[0];
#pragma unroll
  for (int i = 1; i <= 6; i++) q = fma(q, r, kExpC[i]);
  p = q * r;
  double Tj = T[k & 15];
  Ts = __hiloint2double(__double2hiint(Tj) - ((k >> 4) << 20), __double2loint(Tj));
}

// Table-free variant (T == nullptr at compile time is not needed: chosen by the EXPV template parameter):
// exp(-tau) = 2^n * exp(r), n = round(-tau/ln2), |r| <= ln2/2, Taylor through r^12.  ~20 FP64 instructions.
static __constant__ double kExpP[16] = {
    2.08767569878680989792e-09, 2.50521083854417187751e-08, 2.75573192239858906526e-07, 2.75573192239858906526e-06,
    2.48015873015873015873e-05, 1.98412698412698412698e-04, 1.38888888888888888889e-03, 8.33333333333333333333e-03,
    4.16666666666666666667e-02, 1.66666666666666666667e-01, 0.5,
    -1.4426950408889634074,       // [11] -log2(e)
    6755399441055744.0,           // [12]
    -6.93147180369123816490e-01,  // [13] -ln2 hi
    -1.90821492927058770002e-10,  // [14] -ln2 lo
    707.0};
__device__ __forceinline__ void exp_neg_parts_poly(double tau, double& Ts, double& p) {
  tau = __hiloint2double(min(__double2hiint(tau), 0x40861800), __double2loint(tau));
  double t = fma(tau, kExpP[11], kExpP[12]);
  int n = __double2loint(t);
  double fn = t - kExpP[12];
  double r = fma(fn, kExpP[13], -tau);
  r = fma(fn, kExpP[14], r);
  double q = kExpP[0];
#pragma unroll
  for (int i = 1; i <= 10; i++) q = fma(q, r, kExpP[i]);
  p = fma(r * r, q, r);                                       // exp(r) - 1
  Ts = __hiloint2double(0x3ff00000 + (n << 20), 0);           // 2^n
}

template <int EXPV>
__device__ __forceinline__ void exp_neg(double tau, const double* __restrict__ T, double& e, double& ome) {
  double Ts, p;
  if (EXPV == 1) exp_neg_parts(tau, T, Ts, p);
  else exp_neg_parts_poly(tau, Ts, p);
  e = fma(Ts, p, Ts);
  ome = fma(-Ts, p, 1.0 - Ts);
}

template <int EXPV>
__device__ __forceinline__ double exp_neg_only(double tau, const double* __restrict__ T) {
  double Ts, p;
  if (EXPV == 1) exp_neg_parts(tau, T, Ts, p);
  else exp_neg_parts_poly(tau, Ts, p);
  return fma(Ts, p, Ts);
}

// FAST-mode segment: Iout = Iin e^-tau and the segment's contribution to the cell's mean intensity,
//   J_seg * weight/nseg = Iin (1 - e^-tau) / (kappa d) * wn = [Iin (1 - e^-tau)] * cs * (2^-200 / kappa),
// cs = 2^200 wn / d a per-(layer, direction, segment) constant from the host and 2^-200/kappa a per-cell constant that
// the caller applies ONCE per layer to the sum A over all directions and segments (the power of two keeps every
// intermediate far from the subnormal range, including the kappa -> 0 limit).  With Ts = 2^(-k/16), p = e^-r - 1:
//   X = Iin Ts,  Iout = X + X p,  Iin (1 - e^-tau) = (Iin - X) - X p      (Iin - X is exact for k = 0: no cancellation)
// 5 FP64 instructions after the exponential's 12, instead of 8.
// GUARD = false (tau <= 64 for every segment of the warp): no clamp, and no test for Iout underflowing to 0 -- with
// tau <= 64 that can only happen where Iin < 5e-324 e^64 = 3e-296, i.e. where the segment's contribution to Jmean
// (< Iin) is itself below every tolerance floor (the parity tests compare with a floor of 1e-290).
template <int EXPV, bool GUARD = true>
__device__ __forceinline__ double segment_fast(double Iin, double tau, double cs, const double* __restrict__ T, double& A) {
  double Ts, p;
  if (EXPV == 1) exp_neg_parts<GUARD>(tau, T, Ts, p);
  else exp_neg_parts_poly(tau, Ts, p);
  const double X = Iin * Ts;
  const double Iout = fma(X, p, X);
  const double Z = fma(-X, p, Iin - X);
  // Iout == 0 (underflow): the reference gets (Iin - 0)/log(Iin/0) = 0.  Integer test + predicated DFMA.
  if (!GUARD || ((__double2hiint(Iout) << 1) | __double2loint(Iout)) != 0) A = fma(Z, cs, A);
  return Iout;
}

template <int EXPV, bool GUARD = true>
__device__ __forceinline__ double attenuate_fast(double Iin, double tau, const double* __restrict__ T) {
  double Ts, p;
  if (EXPV == 1) exp_neg_parts<GUARD>(tau, T, Ts, p);
  else exp_neg_parts_poly(tau, Ts, p);
  const double X = Iin * Ts;
  return fma(X, p, X);
}

struct SegResult {
  double Iout, J;
};

// FAST mode expects kappa > 0 (callers replace an exact zero by a tiny positive number, for which every formula
// below returns the kappa = 0 limits Iout = Iin, J = Iin) and invtau = 1 / (kappa * dpath).
template <bool FAITHFUL, int EXPV>
__device__ __forceinline__ SegResult segment_update(double Iin, double kappa, double dpath, double invtau,
                                                    const double* __restrict__ T) {
  SegResult r;
  if (FAITHFUL) {
    double tau = __dmul_rn(kappa, dpath);
    double a = exp(-tau);
    r.Iout = __dmul_rn(Iin, a);
    if (r.Iout < Iin) r.J = __ddiv_rn(__dsub_rn(Iin, r.Iout), log(__ddiv_rn(Iin, r.Iout)));
    else r.J = __dmul_rn(0.5, __dadd_rn(Iin, r.Iout));
  } else {
    double tau = kappa * dpath, e, ome;
    exp_neg<EXPV>(tau, T, e, ome);
    r.Iout = Iin * e;
    // Iout == 0 (underflow): the reference gets (Iin - 0)/log(Iin/0) = 0.  (Where Iout is a non-zero SUBNORMAL,
    // i.e. per-segment tau of ~650-745, the reference's log sees only the few bits Iout has left; FAST mode returns
    // the smooth value there -- use FAITHFUL mode to reproduce that artefact.)
    r.J = ((__double_as_longlong(r.Iout) << 1) == 0) ? 0.0 : Iin * (ome * invtau);
  }
  return r;
}

}  // namespace rtb
