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

// The sweeps' exponential (round 2).  The uniform sweep runs into the board's power cap, so its time follows the fp64
// instruction count: a larger table shortens the polynomial.  64 entries: |r| <= ln2/128 = 0.0054, p = e^-r - 1 through
// r^4 -- 3 of the 12 FP64 instructions less per exponential.  Truncation: |dp / p| <= r^4 / 120 = 7.2e-12, i.e. 3.9e-14
// of e^-tau per segment, < 2e-11 after the 512 segments of the longest ray, and 7.2e-12 of a segment's contribution to
// J: two orders of magnitude inside the 1e-9 that FAST arithmetic is held to (measured FAST vs FAITHFUL on the full
// 256^3 x 192 solve: 4.1e-10, unchanged -- that difference is the reference formula's own rounding noise).
// Measured at 256^3 on one box, 20 solves back to back under the power cap: 16 entries / r^7 53.8 ms, 32 / r^5
// 50.7 ms, 64 / r^4 49.8 ms (the 4-way bank conflicts of the larger table cost less than the DFMA saves).
// (The point-source kernels keep the 16-entry / r^7 version above: their deposits are differences of exponentials.)
#ifndef RTB_EXP_TABLE
#define RTB_EXP_TABLE 64
#endif
constexpr int kExpTableSize = RTB_EXP_TABLE;
#if RTB_EXP_TABLE == 32   // 32 entries, |r| <= ln2/64, polynomial through r^5
static __constant__ double kExpC32[12] = {
    -8.33333333333333333333e-03,  // [0] -1/5!
    4.16666666666666666667e-02,   // [1]  1/4!
    -1.66666666666666666667e-01,  // [2] -1/3!
    0.5,                          // [3]
    -1.0,                         // [4]
    46.16624130844683,            // [5]  32/ln2
    6755399441055744.0,           // [6]  1.5 * 2^52
    -0.02166084898635745,         // [7]  -ln2/32, high part (28 trailing zero bits)
    -4.06140840434059e-10,        // [8]  -ln2/32, low part
    0., 0., 0.};
constexpr int kExpDeg = 4;       // Horner steps after the leading coefficient
constexpr int kExpShift = 5;
static __constant__ double kExpTable32[32] = {
    1.00000000000000000000e+00, 9.78572062087700089705e-01, 9.57603280698573700036e-01, 9.37083817055149981279e-01,
    9.17004043204671215328e-01, 8.97354537501553584100e-01, 8.78126080186649726755e-01, 8.59309649061238967072e-01,
    8.40896415253714502036e-01, 8.22877739076982472888e-01, 8.05245165974627141736e-01, 7.87990422553943248296e-01,
    7.71105412703970372057e-01, 7.54582213796711420706e-01, 7.38413072969749673113e-01, 7.22590403488523325137e-01,
    7.07106781186547572737e-01, 6.91954940981916011289e-01, 6.77127773468446325644e-01, 6.62618321579870661608e-01,
    6.48419777325504820276e-01, 6.34525478595866609943e-01, 6.20928906036742001007e-01, 6.07623679990234477621e-01,
    5.94603557501360513449e-01, 5.81862429388788737761e-01, 5.69394317378345782288e-01, 5.57193371297946216103e-01,
    5.45253866332628844837e-01, 5.33570200338411848584e-01, 5.22136891213706877402e-01, 5.10948574327058313571e-01};
#else   // 64 entries, |r| <= ln2/128, polynomial through r^4 (default)
static __constant__ double kExpC32[12] = {
    0., 4.16666666666666666667e-02, -1.66666666666666666667e-01, 0.5, -1.0,
    92.33248261689366,            // [5]  64/ln2
    6755399441055744.0,
    -0.010830424493178725,        // [7]  -ln2/64 hi
    -2.030704202170295e-10,       // [8]  -ln2/64 lo
    0., 0., 0.};
constexpr int kExpDeg = 3;
constexpr int kExpShift = 6;
static __constant__ double kExpTable32[64] = {
    1.00000000000000000000e+00, 9.89228013193975463935e-01, 9.78572062087700089705e-01, 9.68030896746147173637e-01,
    9.57603280698573700036e-01, 9.47287990793482803653e-01, 9.37083817055149981279e-01, 9.26989562541692735387e-01,
    9.17004043204671215328e-01, 9.07126087750199427973e-01, 8.97354537501553584100e-01, 8.87688246263260594127e-01,
    8.78126080186649726755e-01, 8.68666917636853108675e-01, 8.59309649061238967072e-01, 8.50053176859261738763e-01,
    8.40896415253714502036e-01, 8.31838290163368188068e-01, 8.22877739076982472888e-01, 8.14013710928673916989e-01,
    8.05245165974627141736e-01, 7.96571075671133499441e-01, 7.87990422553943248296e-01, 7.79502200118918464611e-01,
    7.71105412703970372057e-01, 7.62799075372269208550e-01, 7.54582213796711420706e-01, 7.46453864145632417504e-01,
    7.38413072969749673113e-01, 7.30458897090323522328e-01, 7.22590403488523325137e-01, 7.14806669195985011633e-01,
    7.07106781186547572737e-01, 6.99489836269155618176e-01, 6.91954940981916011289e-01, 6.84501211487295257996e-01,
    6.77127773468446325644e-01, 6.69833762026651458044e-01, 6.62618321579870661608e-01, 6.55480605762382206869e-01,
    6.48419777325504820276e-01, 6.41435008039389131795e-01, 6.34525478595866609943e-01, 6.27690378512345548145e-01,
    6.20928906036742001007e-01, 6.14240268053435012341e-01, 6.07623679990234477621e-01, 6.01078365726351537823e-01,
    5.94603557501360513449e-01, 5.88198495825140610371e-01, 5.81862429388788737761e-01, 5.75594614976491336655e-01,
    5.69394317378345782288e-01, 5.63260809304120924068e-01, 5.57193371297946216103e-01, 5.51191291653920445448e-01,
    5.45253866332628844837e-01, 5.39380398878559930154e-01, 5.33570200338411848584e-01, 5.27822589180278578525e-01,
    5.22136891213706877402e-01, 5.16512439510614207450e-01, 5.10948574327058313571e-01, 5.05444643025850237628e-01};
#endif

template <bool GUARD = true>
__device__ __forceinline__ void exp_neg_parts32(double tau, const double* __restrict__ T, double& Ts, double& p) {
  if (GUARD) tau = __hiloint2double(min(__double2hiint(tau), 0x40861800), __double2loint(tau));
  double t = fma(tau, kExpC32[5], kExpC32[6]);
  int k = __double2loint(t);
  double fn = t - kExpC32[6];
  double r = fma(fn, kExpC32[7], tau);
  r = fma(fn, kExpC32[8], r);
  double q = kExpC32[4 - kExpDeg];
#pragma unroll
  for (int i = 5 - kExpDeg; i <= 4; i++) q = fma(q, r, kExpC32[i]);
  p = q * r;
  double Tj = T[k & (kExpTableSize - 1)];
  Ts = __hiloint2double(__double2hiint(Tj) - ((k >> kExpShift) << 20), __double2loint(Tj));
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
  if (EXPV == 1) exp_neg_parts32<GUARD>(tau, T, Ts, p);
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
  if (EXPV == 1) exp_neg_parts32<GUARD>(tau, T, Ts, p);
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
