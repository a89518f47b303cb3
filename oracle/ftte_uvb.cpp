// TEST INFRASTRUCTURE ONLY (see ftte_common.h; PARITY UNPINNED: the reference ships no golden vectors and cannot be
// compiled here).  CPU restatement of the UV-background set-up that feeds the diffuse sweep's opacities and the
// diffuse photo-rates:
//   * amplitude model + band amplitudes           equiSources.f90:198-246
//   * powerSpectrumIndex / opposite               equiSources.f90:4985-5058
//   * uvbBetaTable                                uvbBetaTable.f90:1-307
// Fortran real literals without a d-exponent are single precision and are widened when they meet a double
// (SURVEY.md appendix B); `x**n` with an integer n is a repeated product.
#include "ftte_common.h"

namespace ftte {

static inline double f(float x) { return (double)x; }
static inline double ipow2(double x) { return x * x; }
static inline double ipow3(double x) { return x * x * x; }
static inline double ipow4(double x) { const double x2 = x * x; return x2 * x2; }

// equiSources.f90:5046-5058
static bool opposite(double a, double b) { return ((a > 0.) && (b < 0.)) || ((a < 0.) && (b > 0.)); }

// equiSources.f90:4985-5043.  Returns non-zero where the reference prints 'wrong sign' and stops (:5018-5021).
int powerSpectrumIndex(double uvb1, double alpha1, double uvb2, double alpha2, double& uvbTotal, double& alphaTotal,
                       double nug, double nugplus, bool bound) {
  double tmp, tmp1, tmp2, tmpold, fun, fun1, fun2, funToMatch;
  uvbTotal = uvb1 + uvb2;                                                                     // :4997
  if (bound) {
    funToMatch = uvb1 / (alpha1 - 1.) * (1. - pow(nug / nugplus, alpha1 - 1.)) +
                 uvb2 / (alpha2 - 1.) * (1. - pow(nug / nugplus, alpha2 - 1.));               // :5000-5001
  } else {
    funToMatch = uvb1 / (alpha1 - 1.) + uvb2 / (alpha2 - 1.);                                 // :5003
  }
  tmp1 = f(1.1f) * alpha1 - f(0.1f) * alpha2;                                                 // :5006
  tmp2 = f(1.1f) * alpha2 - f(0.1f) * alpha1;                                                 // :5007
  if (bound) {
    fun1 = uvbTotal / (tmp1 - 1.) * (1. - pow(nug / nugplus, tmp1 - 1.)) - funToMatch;        // :5009
    fun2 = uvbTotal / (tmp2 - 1.) * (1. - pow(nug / nugplus, tmp2 - 1.)) - funToMatch;        // :5010
  } else {
    fun1 = uvbTotal / (tmp1 - 1.) - funToMatch;                                               // :5012
    fun2 = uvbTotal / (tmp2 - 1.) - funToMatch;                                               // :5013
  }
  if (!opposite(fun1, fun2)) return ERR_ARG;                                                  // :5016-5019
  tmpold = tmp1;
  tmp = tmp2;
  while (fabs(tmp - tmpold) >= f(1e-8f)) {                                                    // :5023
    tmpold = tmp;
    tmp = (tmp1 * fabs(fun2) + tmp2 * fabs(fun1)) / (fabs(fun1) + fabs(fun2));                // :5025
    if (bound) fun = uvbTotal / (tmp - 1.) * (1. - pow(nug / nugplus, tmp - 1.)) - funToMatch;  // :5027
    else fun = uvbTotal / (tmp - 1.) - funToMatch;                                            // :5029
    if (opposite(fun, fun1)) { tmp2 = tmp; fun2 = fun; }                                      // :5031-5033
    else { tmp1 = tmp; fun1 = fun; }                                                          // :5035-5036
  }
  alphaTotal = tmp;                                                                           // :5040
  return OK;
}

// equiSources.f90:198-246 with contributionQuasar = contributionStellar = 1 (:63-64), alphaQuasar = 1.8,
// alphaStellar = 5. (:61-62).  out: uvb[3], alpha[3], then uvbStellar1..3, uvbQuasar1..3, uniformQuasar, uniformStellar
int uvbAmplitudes(double currentRedshift, double uvbCoefficient, double* uvb, double* alpha, double* extra) {
  const double alphaQuasar = f(1.8f), alphaStellar = f(5.f);
  const double contributionQuasar = 1., contributionStellar = 1.;
  const double z = currentRedshift;
  const double stellar99 = 1. / (1. + ipow4(7. / (1. + z))) * exp(-ipow3(z / 4.));                             // :198
  const double pascal02 = f(0.0188f) * exp(-ipow2(z - 0.5) / (1. + f(0.0625f) * pow(z + f(2.09f), f(2.075f)))) *
                          pow(1. + z, f(3.35f));                                                                 // :206
  const double component1 = stellar99, component2 = pascal02;
  double step = 0.5 * (tanh((z - f(4.2f)) * 1.5) + 1.);                                                          // :212
  const double stellar02 = (1. - step) * component1 + step * component2;                                         // :213
  const double quasar02 = 10. / (1. + ipow4(7. / (1. + z))) * exp(-ipow3(z / 2.5));                              // :217
  const double gaussian = exp(-ipow2((z - 4.5) / 2.)) * f(0.3f);                                                 // :219
  const double newQuasar02 = gaussian * stellar02 + (1. - gaussian) * quasar02;                                  // :221
  double newStellar02 = (1. - gaussian) * stellar02 + gaussian * quasar02;                                       // :222
  step = 0.5 * (tanh((z - 14.) * 0.5) + 1.);                                                                     // :224
  newStellar02 = step * 0. + (1. - step) * newStellar02;                                                         // :225
  const double uniformQuasar = newQuasar02 * 1.e-21 * contributionQuasar * uvbCoefficient;                       // :231
  const double uniformStellar = newStellar02 * 1.e-21 * contributionStellar * uvbCoefficient;                    // :232
  const double uvbStellar1 = newStellar02 * 1.e-21 * contributionStellar * uvbCoefficient;                       // :236
  const double uvbStellar2 = uvbStellar1 * pow(nu2 / nu1, -alphaStellar);                                        // :237
  const double uvbStellar3 = uvbStellar2 * pow(nu3 / nu2, -alphaStellar);                                        // :238
  const double uvbQuasar1 = newQuasar02 * 1.e-21 * contributionQuasar * uvbCoefficient;                          // :240
  const double uvbQuasar2 = uvbQuasar1 * pow(nu2 / nu1, -alphaQuasar);                                           // :241
  const double uvbQuasar3 = uvbQuasar2 * pow(nu3 / nu2, -alphaQuasar);                                           // :242
  int st = powerSpectrumIndex(uvbStellar1, alphaStellar, uvbQuasar1, alphaQuasar, uvb[0], alpha[0], nu1, nu2, true);   // :244
  if (!st) st = powerSpectrumIndex(uvbStellar2, alphaStellar, uvbQuasar2, alphaQuasar, uvb[1], alpha[1], nu2, nu3, true);   // :245
  if (!st) st = powerSpectrumIndex(uvbStellar3, alphaStellar, uvbQuasar3, alphaQuasar, uvb[2], alpha[2], nu3, nu3, false);  // :246
  if (extra) {
    extra[0] = uvbStellar1; extra[1] = uvbStellar2; extra[2] = uvbStellar3;
    extra[3] = uvbQuasar1; extra[4] = uvbQuasar2; extra[5] = uvbQuasar3;
    extra[6] = uniformQuasar; extra[7] = uniformStellar;
  }
  return st;
}

// uvbBetaTable.f90:1-307.  out = [3 groups][19]: beta24..beta31 (8), ksi24..ksi31 (8), gammaHI, gammaHeI, gammaHeII.
void uvbBetaTable(int nfreq, double freqdel, const double* alpha, double* out) {
  const double e27 = f(0.755f), e28a = f(2.65f), e28b = f(11.27f), e28c = f(21.0f), e29a = f(15.42f),
               e29b = f(16.5f), e29c = f(17.7f), e30a = f(30.0f), e30b = f(70.0f);        // :20-29
  std::vector<double> sigma[8], nu((size_t)nfreq);
  for (auto& s : sigma) s.assign((size_t)nfreq, 0.);
  double* sigma24 = sigma[0].data(); double* sigma25 = sigma[1].data(); double* sigma26 = sigma[2].data();
  double* sigma27 = sigma[3].data(); double* sigma28 = sigma[4].data(); double* sigma29 = sigma[5].data();
  double* sigma30 = sigma[6].data(); double* sigma31 = sigma[7].data();
  for (int i = 1; i <= nfreq; i++) {
    const int q = i - 1;
    nu[q] = pow(10.0, (double)(i - 1) * freqdel);                                                         // :33
    const double x = nu[q];
    double dum;
    if (x > hydrogenIonization) {                                                                         // :35-44
      dum = sqrt(x / hydrogenIonization - 1);
      sigma24[q] = f(6.3e-18f) * ipow4(hydrogenIonization / x) * exp(4.0 - 4.0 * atan(dum) / dum) /
                   (1 - exp(-2.0 * pi / dum));
    }
    if (x > doubleHeliumIonization) {                                                                     // :46-55
      dum = sqrt(x / doubleHeliumIonization - 1);
      sigma25[q] = f(1.58e-18f) * ipow4(doubleHeliumIonization / x) * exp(4.0 - 4.0 * atan(dum) / dum) /
                   (1 - exp(-2.0 * pi / dum));
    }
    if (x > singleHeliumIonization) {                                                                     // :57-65
      sigma26[q] = f(7.42e-18f) * (f(1.66f) * pow(x / singleHeliumIonization, f(-2.05f)) -
                                   f(0.66f) * pow(x / singleHeliumIonization, f(-3.05f)));
    }
    if (x > e27) sigma27[q] = f(2.11e-16f) * pow(x - e27, 1.5) / ipow3(x);                                // :67-71
    if (x > e28a && x <= e28b)                                                                            // :73-79
      sigma28[q] = pow(10.0, f(-40.97f) + f(6.03f) * x - f(0.504f) * ipow2(x) + f(1.387e-2f) * ipow3(x));
    else if (x > e28b && x < e28c)
      sigma28[q] = pow(10.0, f(-30.26f) + f(2.79f) * x - f(0.184f) * ipow2(x) + f(3.535e-3f) * ipow3(x));
    if (x > e29a && x <= e29b) sigma29[q] = f(6.2e-18f) * x - f(9.4e-17f);                                // :81-89
    else if (x > e29b && x <= e29c) sigma29[q] = f(1.4e-18f) * x - f(1.48e-17f);
    else if (x > e29c) sigma29[q] = f(2.5e-14f) * pow(x, f(-2.71f));
    if (x >= e30a && x < e30b)                                                                            // :91-95
      sigma30[q] = pow(10.0, f(-16.926f) - f(4.528e-2f) * x + f(2.238e-4f) * ipow2(x) + f(4.245e-7f) * ipow3(x));
    if (x > e28b && x < hydrogenIonization) sigma31[q] = f(3.71e-18f);                                    // :97-101
  }
  for (int i = 0; i < 57; i++) out[i] = 0.;                                                               // :111-169
  double* g1 = out; double* g2 = out + 19; double* g3 = out + 38;
  for (int i = 2; i <= nfreq; i++) {                                                                      // :171
    const int q = i - 1;
    const double freq = nu[q];
    const double delta_nu = nu[q] - nu[q - 1];
    if (freq >= nu1 && freq <= nu2) {                                                                     // :176-200
      const double dtmp = pow(freq / nu1, -alpha[0]) * delta_nu;
      const double dtmpOverEnergy = dtmp * eV_to_Hz / (freq * eV_to_erg);
      for (int s = 0; s < 8; s++) g1[s] = g1[s] + dtmp * sigma[s][q];
      for (int s = 0; s < 8; s++) g1[8 + s] = g1[8 + s] + dtmpOverEnergy * sigma[s][q];
      g1[16] = g1[16] + dtmpOverEnergy * (freq - nu1) * eV_to_erg * sigma24[q];
    }
    if (freq >= nu2 && freq <= nu3) {                                                                     // :202-227
      const double dtmp = pow(freq / nu2, -alpha[1]) * delta_nu;
      const double dtmpOverEnergy = dtmp * eV_to_Hz / (freq * eV_to_erg);
      for (int s = 0; s < 8; s++) g2[s] = g2[s] + dtmp * sigma[s][q];
      for (int s = 0; s < 8; s++) g2[8 + s] = g2[8 + s] + dtmpOverEnergy * sigma[s][q];
      g2[16] = g2[16] + dtmpOverEnergy * (freq - nu1) * eV_to_erg * sigma24[q];
      g2[17] = g2[17] + dtmpOverEnergy * (freq - nu2) * eV_to_erg * sigma26[q];
    }
    if (freq >= nu3) {                                                                                    // :229-255
      const double dtmp = pow(freq / nu3, -alpha[2]) * delta_nu;
      const double dtmpOverEnergy = dtmp * eV_to_Hz / (freq * eV_to_erg);
      for (int s = 0; s < 8; s++) g3[s] = g3[s] + dtmp * sigma[s][q];
      for (int s = 0; s < 8; s++) g3[8 + s] = g3[8 + s] + dtmpOverEnergy * sigma[s][q];
      g3[16] = g3[16] + dtmpOverEnergy * (freq - nu1) * eV_to_erg * sigma24[q];
      g3[17] = g3[17] + dtmpOverEnergy * (freq - nu2) * eV_to_erg * sigma26[q];
      g3[18] = g3[18] + dtmpOverEnergy * (freq - nu3) * eV_to_erg * sigma25[q];
    }
  }
  const double shape1 = (1. - pow(nu2 / nu1, 1. - alpha[0])) / (alpha[0] - 1.);                           // :259
  const double shape2 = (1. - pow(nu3 / nu2, 1. - alpha[1])) / (alpha[1] - 1.);                           // :260
  const double shape3 = 1. / (alpha[2] - 1.);                                                             // :261
  const double energyShape1 = shape1 * nu1, energyShape2 = shape2 * nu2, energyShape3 = shape3 * nu3;
  for (int s = 0; s < 8; s++) g1[s] = g1[s] / energyShape1;                                               // :265-273
  for (int s = 0; s < 8; s++) g2[s] = g2[s] / energyShape2;                                               // :275-283
  for (int s = 0; s < 8; s++) g3[s] = g3[s] / energyShape3;                                               // :285-293
}

}  // namespace ftte

extern "C" {
// out57 = [group 1..3][beta24, beta25, beta26, beta27..31, ksi24, ksi25, ksi26, ksi27..31, gammaHI, gammaHeI, gammaHeII]
int ftte_uvb(double currentRedshift, double uvbCoefficient, int nfreq, double freqdel, double* uvb, double* alpha,
             double* extra, double* out57) {
  int st = ftte::uvbAmplitudes(currentRedshift, uvbCoefficient, uvb, alpha, extra);
  if (st) return st;
  ftte::uvbBetaTable(nfreq, freqdel, alpha, out57);
  return ftte::OK;
}
int ftte_power_spectrum_index(double uvb1, double alpha1, double uvb2, double alpha2, double nug, double nugplus,
                              int bound, double* uvbTotal, double* alphaTotal) {
  return ftte::powerSpectrumIndex(uvb1, alpha1, uvb2, alpha2, *uvbTotal, *alphaTotal, nug, nugplus, bound != 0);
}
void ftte_uvb_beta_table(int nfreq, double freqdel, const double* alpha, double* out57) {
  ftte::uvbBetaTable(nfreq, freqdel, alpha, out57);
}
}
