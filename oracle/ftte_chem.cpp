// TEST INFRASTRUCTURE ONLY (see ftte_common.h).  PARITY UNPINNED (no reference golden vectors exist).
//
// CPU restatement of solveRateEquations (equiSources.f90:3459-3677): per leaf, photo-rates per absorber from the
// per-cell point-source rates (:3518-3543), the diffuse (:3546-3553) or uniform (:3555-3562) background contribution,
// linear lookup of k1..k6 in log T (:3566-3586), bisection on the electron density for H/He ionisation equilibrium
// (:3590-3631), new HI, HeI, HeII (:3633-3673).  The rate-coefficient tables k1a..k6a come from calc_rates.f
// (out of scope) and are inputs.  Single thread, libm log, no FMA contraction.
#include "ftte_common.h"

namespace ftte {

static inline bool opposite(double a, double b) {  // equiSources.f90:5044-5058
  return ((a > 0.) && (b < 0.)) || ((a < 0.) && (b > 0.));
}

// status: 0 ok, 15 = ionisation fraction out of range (the reference prints and stops, :3638-3655)
int chemistrySolve(int64_t nleaf, int nx, double physicalBoxSize, const int8_t* level, const double* rho,
                   const double* tgas, double* HI_, double* HeI_, double* HeII_, const double* rates /* [6][nleaf] or null */,
                   const double* J /* [3][nleaf] or null: uniform background */, const double* ksi /* ksi24[3], ksi25, ksi26[2] */,
                   const double* uniform /* add24, add25, add26, selfShieldingThreshold */, int nratec, double logtem0,
                   double logtem9, double dlogtem, const double* k1a, const double* k2a, const double* k3a,
                   const double* k4a, const double* k5a, const double* k6a, double* maxChange) {
  double worst = 0.;
  for (int64_t c = 0; c < nleaf; c++) {
    const double nh = psi * rho[c] / mh;
    const double nhe = (1. - psi) * rho[c] / mhe;
    double HI = std::fmin(HI_[c], nh);
    double HII = nh - HI_[c];
    double HeI = HeI_[c];
    double HeII = HeII_[c];
    double HeIII = nhe - HeI_[c] - HeII_[c];
    double cellHI = HI_[c], cellHeI = HeI_[c], cellHeII = HeII_[c];
    if (HeIII < 0.) {
      cellHeII = nhe - cellHeI;
      HeIII = 0.;
      if (HeII < 0.) { cellHeI = nhe; cellHeII = 0.; HeII = 0.; HeIII = 0.; }
    }
    double de = HII + HeII + 2. * HeIII;
    const double physicalCellSize = physicalBoxSize / ((double)(float)(1 << level[c]) * (double)(float)nx);
    const double vol = physicalCellSize * physicalCellSize * physicalCellSize;
    double krate24 = 0., krate25 = 0., krate26 = 0.;
    if (rates) {
      if (HI > 0.) krate24 = rates[0 * nleaf + c] / (vol * HI);
      if (HeII > 0.) krate25 = rates[1 * nleaf + c] / (vol * HeII);
      if (HeI > 0.) krate26 = rates[2 * nleaf + c] / (vol * HeI);
    }
    krate24 = std::fmax(krate24, 0.);
    krate25 = std::fmax(krate25, 0.);
    krate26 = std::fmax(krate26, 0.);
    if (J) {
      const double tmp1 = 4. * pi * J[0 * nleaf + c], tmp2 = 4. * pi * J[1 * nleaf + c], tmp3 = 4. * pi * J[2 * nleaf + c];
      krate24 = krate24 + tmp1 * ksi[0] + tmp2 * ksi[1] + tmp3 * ksi[2];
      krate25 = krate25 + tmp3 * ksi[3];
      krate26 = krate26 + tmp2 * ksi[4] + tmp3 * ksi[5];
    } else {
      const double mfp = 1. / (HI * (double)6.3e-18f + HeI * (double)7.42e-18f + HeII * (double)1.58e-18f);
      if (mfp >= uniform[3]) {
        krate24 = krate24 + uniform[0];
        krate25 = krate25 + uniform[1];
        krate26 = krate26 + uniform[2];
      }
    }
    de = HII + HeII + 2. * HeIII;
    double logtem = std::log(tgas[c]);
    logtem = std::fmax(logtem, logtem0);
    logtem = std::fmin(logtem, logtem9);
    const int indixe = std::min(nratec - 1, std::max(1, (int)((logtem - logtem0) / dlogtem) + 1));
    const double t1 = logtem0 + (indixe - 1) * dlogtem, t2 = logtem0 + indixe * dlogtem, tdef = t2 - t1;
    auto look = [&](const double* k) { return k[indixe - 1] + (logtem - t1) * (k[indixe] - k[indixe - 1]) / tdef; };
    const double k1 = look(k1a), k2 = look(k2a), k3 = look(k3a), k4 = look(k4a), k5 = look(k5a), k6 = look(k6a);
    auto heI = [&](double d) {
      return (d - nh / (1. + k2 * d / (k1 * d + krate24)) - 2. * nhe) /
             ((k3 * d + krate26) / (k4 * d) - 2. - 2. * (k3 * d + krate26) / (k4 * d));
    };
    auto resid = [&](double h, double d) {
      return k3 * h * d + k6 * (nhe - h - h * (k3 * d + krate26) / (k4 * d)) * d + krate26 * h -
             h * (k3 * d + krate26) / (k4 * d) * (k4 * d + k5 * d + krate25);
    };
    double de1 = (double)1.e-30f;
    de = de1;
    HeI = heI(de);
    double res1 = resid(HeI, de);
    double de2 = nh + 2. * nhe;
    de = de2;
    HeI = heI(de);
    double res2 = resid(HeI, de);
    double HeIprev = -1.;
    int guard = 0;
    while (std::fabs(HeI - HeIprev) / nhe > 1.e-10) {
      HeIprev = HeI;
      de = 0.5 * (de1 + de2);
      HeI = heI(de);
      const double res = resid(HeI, de);
      if (opposite(res, res1)) { de2 = de; res2 = res; }
      else { de1 = de; res1 = res; }
      if (++guard > 100000) return 15;
    }
    (void)res2;
    HeII = HeI * (k3 * de + krate26) / (k4 * de);
    HeIII = nhe - HeI - HeII;
    HII = nh / (1. + k2 * de / (k1 * de + krate24));
    HI = k2 * HII * de / (k1 * de + krate24);
    if (!(HI / nh >= 0. && HI / nh <= 1.)) return 15;
    if (!(HeI / nhe >= 0. && HeI / nhe <= 1.)) return 15;
    const double tmp1 = std::fabs(HI - cellHI) * mh / (psi * rho[c]);
    const double tmp2 = std::fabs(HeI - cellHeI) * mhe / ((1. - psi) * rho[c]);
    const double tmp3 = std::fabs(HeII - cellHeII) * mhe / ((1. - psi) * rho[c]);
    worst = std::fmax(worst, std::fmax(tmp1, std::fmax(tmp2, tmp3)));
    HI_[c] = HI; HeI_[c] = HeI; HeII_[c] = HeII;
  }
  if (maxChange) *maxChange = worst;
  return OK;
}

}  // namespace ftte
