// TEST INFRASTRUCTURE ONLY (see ftte_common.h).  PARITY UNPINNED (no reference golden vectors exist).
//
// CPU restatement of the point-source path of razoumov/radiativeTransfer:
//   equiSources.f90:293-309 (rmax), :1256-1370 (source loop), :2412-2595 (drawSegment), :2647-2960 (find/zoom
//   neighbours), :3011-3118 (absoluteCoordinates, localizeSplitContinuationCell), :3120-3385 (startNewLongRay),
//   :4157-4311 (getRatesHydrogenHelium); stellarBetaTable.f90:31-285; stellarPopulationModule.f90:7-50;
//   dustModule.f90:30-73.
// Single thread, libm, no FMA contraction.
#include <algorithm>

#include "ftte_common.h"
// exp/log built from IEEE +,*,/,fma only; lives with the product because the device kernels use the same source.
// With ftte_set_portable_math(1) the oracle evaluates the table sums, the table lookups and the escape diagnostics
// with it instead of libm, which makes the CUDA path (RTB200_MATH_FAITHFUL) reproduce every deposit bit for bit.
// Default is libm, as gfortran would link.
#include "../radiativetransfer_b200/csrc/portable_math.h"

namespace ftte {

static bool g_portableMath = false;
void setPortableMath(int on) { g_portableMath = on != 0; }
static inline double xexp(double x) { return g_portableMath ? rtb_pm::pm_exp(x) : std::exp(x); }
static inline double xlog(double x) { return g_portableMath ? rtb_pm::pm_log(x) : std::log(x); }

static const int ndepth = 10;  // ndepth1..3, ndepthDust (definitionsModule.f90:72)
static const int nT = 11 * 11 * 11 * 11;
static inline int tix(int i1, int i2, int i3, int iD) { return ((iD * 11 + i3) * 11 + i2) * 11 + i1; }  // Fortran order

struct PointTables {
  std::vector<double> R[3], E[3];  // reactionRate1..3, energyRate1..3, each (0:10)^4
  double totalIntegral;
  double outputFreq[300], outputSigma24[300], outputSigma25[300], outputSigma26[300], outputSigmaDust[300];
};

// dustModule.f90:30-73 (SMC = 1)
static double dustCrossSection(double lambda, const double* a /* [7][5] row-major: a(i,1..5) */) {
  double sigma = 0;
  for (int i = 0; i < 7; i++) {
    double x = lambda / a[i * 5 + 0];
    sigma = sigma + a[i * 5 + 1] / (std::pow(x, a[i * 5 + 3]) + std::pow(x, -a[i * 5 + 4]) + a[i * 5 + 2]);
  }
  return (double)1.1f * sigma * (double)0.9210340372f;
}

// stellarPopulationModule.f90:7-50; lum = [nMet][2][nWave] (the two time slices iSpectrum, iSpectrum+1)
static double stellarPopulation(const PointSpectra& S, double coefSpectrum, int iMetal, double coefMetal, double freq) {
  double thisWavelength = clight / (freq * eV_to_Hz);
  int iW = 1;
  while (thisWavelength > S.wavelength[iW]) iW++;  // wavelength(iWavelength+1), 1-based -> [iW]
  double coefW = (thisWavelength - S.wavelength[iW - 1]) / (S.wavelength[iW] - S.wavelength[iW - 1]);
  coefW = std::fmin(std::fmax(0., coefW), 1.);
  auto L = [&](int m, int t, int w) { return S.lum[((size_t)(m - 1) * 2 + t) * S.nWave + (w - 1)]; };
  double sp1 = coefSpectrum * ((1. - coefW) * L(iMetal, 1, iW) + coefW * 1. * L(iMetal, 1, iW + 1)) +
               (1. - coefSpectrum) * ((1. - coefW) * L(iMetal, 0, iW) + coefW * L(iMetal, 0, iW + 1));
  double sp2 = coefSpectrum * ((1. - coefW) * L(iMetal + 1, 1, iW) + coefW * 1. * L(iMetal + 1, 1, iW + 1)) +
               (1. - coefSpectrum) * ((1. - coefW) * L(iMetal + 1, 0, iW) + coefW * L(iMetal + 1, 0, iW + 1));
  double SP = (1. - coefMetal) * sp1 + coefMetal * sp2;
  double f = freq * eV_to_Hz;
  return std::pow(10., SP) / angstrom * clight / (f * f);
}

static inline double pow4(double x) { double y = x * x; return y * y; }  // x**4 by repeated squaring

// stellarBetaTable.f90
void stellarBetaTable(const PointSpectra& S, int iMetal, double coefMetal, PointTables& T) {
  const int nfreq = 400;
  const double freqdel = (double)0.02f;
  std::vector<double> nu(nfreq + 1), s24(nfreq + 1), s25(nfreq + 1), s26(nfreq + 1), sD(nfreq + 1);
  for (int i = 1; i <= nfreq; i++) {
    nu[i] = std::pow(10.0, (double)(i - 1) * freqdel);
    double lambda = clight / (nu[i] * eV_to_Hz) * (double)1.e8f;
    sD[i] = dustCrossSection(lambda / (double)1.e4f, S.aDust) * (double)1.e-22f;
    if (nu[i] > hydrogenIonization) {
      double dum = std::sqrt(nu[i] / hydrogenIonization - 1);
      s24[i] = (double)6.3e-18f * pow4(hydrogenIonization / nu[i]) * std::exp(4.0 - 4.0 * std::atan(dum) / dum) /
               (1 - std::exp(-2.0 * pi / dum));
    } else s24[i] = 0.;
    if (nu[i] > doubleHeliumIonization) {
      double dum = std::sqrt(nu[i] / doubleHeliumIonization - 1);
      s25[i] = (double)1.58e-18f * pow4(doubleHeliumIonization / nu[i]) * std::exp(4.0 - 4.0 * std::atan(dum) / dum) /
               (1 - std::exp(-2.0 * pi / dum));
    } else s25[i] = 0.;
    if (nu[i] > singleHeliumIonization) {
      s26[i] = (double)7.42e-18f * ((double)1.66f * std::pow(nu[i] / singleHeliumIonization, (double)(-2.05f)) -
                                    (double)0.66f * std::pow(nu[i] / singleHeliumIonization, (double)(-3.05f)));
    } else s26[i] = 0.;
  }
  const double lowerEnergy = hydrogenIonization, upperEnergy = 10. * hydrogenIonization;
  for (int ie = 1; ie <= 300; ie++) {
    double freq = lowerEnergy * std::exp((double)((float)(ie - 1) / (float)(300 - 1)) * (std::log(upperEnergy) - std::log(lowerEnergy)));
    T.outputFreq[ie - 1] = freq;
    double lambda = clight / (freq * eV_to_Hz) * (double)1.e8f;
    T.outputSigmaDust[ie - 1] = dustCrossSection(lambda / (double)1.e4f, S.aDust) * (double)1.e-22f;
    if (freq > hydrogenIonization) {
      double dum = std::sqrt(freq / hydrogenIonization - 1);
      T.outputSigma24[ie - 1] = (double)6.3e-18f * pow4(hydrogenIonization / freq) * std::exp(4. - 4. * std::atan(dum) / dum) /
                                (1 - std::exp(-2. * pi / dum));
    } else if (freq == hydrogenIonization) T.outputSigma24[ie - 1] = (double)6.3e-18f;
    else T.outputSigma24[ie - 1] = 0.;
    if (freq > doubleHeliumIonization) {
      double dum = std::sqrt(freq / doubleHeliumIonization - 1);
      T.outputSigma25[ie - 1] = (double)1.58e-18f * pow4(doubleHeliumIonization / freq) * std::exp(4. - 4. * std::atan(dum) / dum) /
                                (1 - std::exp(-2. * pi / dum));
    } else T.outputSigma25[ie - 1] = 0.;
    if (freq > singleHeliumIonization) {
      T.outputSigma26[ie - 1] = (double)7.42e-18f * ((double)1.66f * std::pow(freq / singleHeliumIonization, (double)(-2.05f)) -
                                                     (double)0.66f * std::pow(freq / singleHeliumIonization, (double)(-3.05f)));
    } else T.outputSigma26[ie - 1] = 0.;
  }
  T.totalIntegral = 0.;
  for (int r = 0; r < 3; r++) { T.R[r].assign(nT, 0.); T.E[r].assign(nT, 0.); }
  const double thr[3] = {nu1, nu2, nu3};
  for (int i = 2; i <= nfreq; i++) {
    double freq = nu[i];
    double delta_nu = nu[i] - nu[i - 1];
    double lum = stellarPopulation(S, S.coefSpectrum, iMetal, coefMetal, freq);
    double dtmp = lum / (freq * eV_to_erg) * delta_nu * eV_to_Hz;
    if (freq >= nu1) T.totalIntegral = T.totalIntegral + dtmp;
    for (int i1 = 0; i1 <= ndepth; i1++)
      for (int i2 = 0; i2 <= ndepth; i2++)
        for (int i3 = 0; i3 <= ndepth; i3++)
          for (int iD = 0; iD <= ndepth; iD++) {
            double tau1 = (double)((float)i1 / (float)ndepth) * 10.;
            double tau2 = (double)((float)i2 / (float)ndepth) * 10.;
            double tau3 = (double)((float)i3 / (float)ndepth) * 10.;
            double tauDust = (double)((float)iD / (float)ndepth) * 10.;
            tau1 = s24[i] / (double)6.3e-18f * tau1;
            tau2 = s26[i] / (double)7.42e-18f * tau2;
            tau3 = s25[i] / (double)1.58e-18f * tau3;
            tauDust = sD[i] / (double)5.4116737e-22f * tauDust;
            const int e = tix(i1, i2, i3, iD);
            for (int r = 0; r < 3; r++)
              if (freq >= thr[r]) {
                double atmp = dtmp * xexp(-(tau1 + tau2 + tau3 + tauDust));
                T.R[r][e] = T.R[r][e] + atmp;
                T.E[r][e] = T.E[r][e] + (freq - thr[r]) * eV_to_erg * atmp;
              }
          }
  }
}

// equiSources.f90:4157-4311
static int getRates(const PointTables& T, int dustApproximation, int reaction, double tau1, double tau2, double tau3,
                    double tauDust, double& numberRate, double& heatingRate) {
  if (tau1 > 10. || tau2 > 10. || tau3 > 10. || tauDust > 10.) { numberRate = 0.; heatingRate = 0.; return OK; }
  int id1 = (int)(tau1 / 10. * (double)(float)ndepth);
  int id2 = (int)(tau2 / 10. * (double)(float)ndepth);
  int id3 = (int)(tau3 / 10. * (double)(float)ndepth);
  double c1 = tau1 * (double)(float)ndepth / 10. - (double)(float)id1;
  double c2 = tau2 * (double)(float)ndepth / 10. - (double)(float)id2;
  double c3 = tau3 * (double)(float)ndepth / 10. - (double)(float)id3;
  int idD; double cD;
  if (dustApproximation == 0) { idD = 0; cD = 0.; }
  else {
    idD = (int)(tauDust / 10. * (double)(float)ndepth);
    cD = tauDust * (double)(float)ndepth / 10. - (double)(float)idD;
  }
  if (std::min(std::min(id1, id2), std::min(id3, idD)) < 0) return ERR_IDEPTH;
  if (id1 >= ndepth || id2 >= ndepth || id3 >= ndepth || idD >= ndepth) return ERR_IDEPTH;  // tau == 10 exactly: the reference reads out of bounds
  auto interp = [&](const std::vector<double>& A, int iD) {
    auto lg = [&](int a, int b, int c) { return xlog(A[tix(a, b, c, iD)]); };
    return c1 * ((1. - c3) * (1. - c2) * lg(id1 + 1, id2, id3) + c3 * (1. - c2) * lg(id1 + 1, id2, id3 + 1) +
                 c2 * (1. - c3) * lg(id1 + 1, id2 + 1, id3) + c3 * c2 * lg(id1 + 1, id2 + 1, id3 + 1)) +
           (1. - c1) * ((1. - c3) * (1. - c2) * lg(id1, id2, id3) + c3 * (1. - c2) * lg(id1, id2, id3 + 1) +
                        c2 * (1. - c3) * lg(id1, id2 + 1, id3) + c3 * c2 * lg(id1, id2 + 1, id3 + 1));
  };
  const int r = reaction - 1;
  double nr1 = interp(T.R[r], idD), nr2 = interp(T.R[r], idD + 1);
  numberRate = xexp((1. - cD) * nr1 + cD * nr2);
  double hr1 = interp(T.E[r], idD), hr2 = interp(T.E[r], idD + 1);
  heatingRate = xexp((1. - cD) * hr1 + cD * hr2);
  return OK;
}

struct Pt { double x, y, z; };

struct PointSolver {
  Grid& g;
  const PointTables* T = nullptr;
  int dustApproximation = 0, maxPixelLevel = 6;
  double rmax[31];
  double outputRadius[7];
  // per-source diagnostics (equiSources.f90:9-13)
  double ndotRemaining[7], ndotBoundary[7], ndotDust, ndotSpectrum[300];
  int highestPixelLevel = 0;
  int64_t nseg = 0;
  int64_t* trace = nullptr; int64_t traceCap = 0, traceLen = 0;  // leaf<<32 | pixelLevel<<28 | ipix<<8 | exit face, per segment
  // module globals of the reference (definitionsModule.f90:276-282)
  int neighbourCell = -1; double xneighbour = 0, yneighbour = 0, zneighbour = 0;
  int neighbourSeq[40], neighbourLevel = 0;
  double xbase = 0, ybase = 0, zbase = 0;
  int splitCell = -1, splitSeq[40]; Pt splitPoint;
  explicit PointSolver(Grid& gg) : g(gg) {
    for (int ir = 1; ir <= 30; ir++) {  // equiSources.f90:304-309, single-precision expression
      float v = std::sqrt(3.f) * (std::sqrt(0.5f * std::pow(4.f, (float)(ir - 1)) - 1.f / 12.f) + 0.5f);
      rmax[ir] = (double)v / 2.;
    }
    const float orad[7] = {0.1f, 0.3f, 1.f, 3.f, 10.f, 30.f, 100.f};
    for (int i = 0; i < 7; i++) outputRadius[i] = (double)orad[i];
  }
  int cellAt(const int* seq, int level) const {
    int n = g.base(seq[0], seq[1], seq[2]);
    for (int l = 1; l <= level; l++) n = g.kid(n, seq[3 * l], seq[3 * l + 1], seq[3 * l + 2]);
    return n;
  }

  // zoom??Neighbour (equiSources.f90:2827-2960): axis = the axis normal to the face (0 x [yz plane], 1 y [xz], 2 z [xy])
  void zoom(int cell, int level, int* seq, int axis, double a, double b, int side) {
    while (g.node[cell].refined()) {
      int ia, ib; double an, bn;
      if (a < 0.5) { an = 2. * a; ia = 1; } else { an = 2. * a - 1.; ia = 2; }
      if (b < 0.5) { bn = 2. * b; ib = 1; } else { bn = 2. * b - 1.; ib = 2; }
      int in = side == 0 ? 2 : 1;
      int i, j, k;
      if (axis == 2) { i = ia; j = ib; k = in; }        // (x, y) on an xy face
      else if (axis == 0) { i = in; j = ia; k = ib; }   // (y, z) on a yz face
      else { i = ia; j = in; k = ib; }                  // (x, z) on an xz face
      seq[3 * level + 3] = i; seq[3 * level + 4] = j; seq[3 * level + 5] = k;
      cell = g.kid(cell, i, j, k);
      level++;
      a = an; b = bn;
    }
    neighbourCell = cell;
    neighbourLevel = level;
    if (axis == 2) { xneighbour = a; yneighbour = b; }
    else if (axis == 0) { yneighbour = a; zneighbour = b; }
    else { xneighbour = a; zneighbour = b; }
    std::memcpy(neighbourSeq, seq, sizeof(int) * (3 * level + 3));
  }

  // find??Neighbour (equiSources.f90:2647-2825)
  void findNeighbour(int level, const int* seqIn, int axis, double a, double b, int side, int& strategy) {
    int seq[40];
    std::memcpy(seq, seqIn, sizeof(int) * (3 * level + 3));
    // the two in-face axes, in the order (a, b): xy face -> (x, y); yz face -> (y, z); xz face -> (x, z)
    const int axA = axis == 2 ? 0 : (axis == 0 ? 1 : 0);
    const int axB = axis == 2 ? 1 : 2;
    while (level > 0) {
      const int cur = seq[3 * level + axis];
      if ((side == 0 && cur == 1) || (side == 1 && cur == 2)) {
        a = seq[3 * level + axA] == 1 ? 0.5 * a : 0.5 * a + 0.5;
        b = seq[3 * level + axB] == 1 ? 0.5 * b : 0.5 * b + 0.5;
        level--;
      } else {
        seq[3 * level + axis] = side == 0 ? 1 : 2;
        zoom(cellAt(seq, level), level, seq, axis, a, b, side);
        return;
      }
    }
    const int nmax = axis == 0 ? g.nx : (axis == 1 ? g.ny : g.nz);
    if ((side == 0 && seq[axis] == 1) || (side == 1 && seq[axis] == nmax)) { strategy = boundary; return; }
    seq[axis] = side == 0 ? seq[axis] - 1 : seq[axis] + 1;
    zoom(cellAt(seq, 0), 0, seq, axis, a, b, side);
  }

  // drawSegment (equiSources.f90:2412-2595)
  int drawSegment(int cell, Pt& sp, double phi, double theta, int pixelLevel, int level, const int* seq, double& radius,
                  int& strategy, double& length, int& face) {
    if (g.node[cell].level != level) return ERR_ARG;
    double prox = std::cos(phi) * std::cos(theta);
    double proy = std::sin(phi) * std::cos(theta);
    double proz = std::sin(theta);
    double tmp1 = proz > 0. ? (1. - sp.z) / proz : -sp.z / proz;
    double tmp2 = prox > 0. ? (1. - sp.x) / prox : -sp.x / prox;
    double tmp3 = proy > 0. ? (1. - sp.y) / proy : -sp.y / proy;
    int dir; double tmp;
    if (tmp1 < std::fmin(tmp2, tmp3)) { dir = 1; tmp = tmp1; }       // xyPlane
    else if (tmp2 < std::fmin(tmp1, tmp3)) { dir = 2; tmp = tmp2; }  // yzPlane
    else { dir = 3; tmp = tmp3; }                                    // xzPlane
    face = 0;
    const double scale = (double)(float)(1 << level);
    if ((radius * scale + tmp < rmax[pixelLevel]) || pixelLevel == maxPixelLevel) {
      strategy = proceed;
      length = tmp;
      radius = radius + tmp / scale;
      Pt e{sp.x + tmp * prox, sp.y + tmp * proy, sp.z + tmp * proz};
      int side;
      if (dir == 1) {
        side = proz < 0. ? 0 : 1;
        findNeighbour(level, seq, 2, e.x, e.y, side, strategy);
        if (strategy != boundary) { sp.z = side == 0 ? 1. : 0.; sp.x = xneighbour; sp.y = yneighbour; }
      } else if (dir == 2) {
        side = prox < 0. ? 0 : 1;
        findNeighbour(level, seq, 0, e.y, e.z, side, strategy);
        if (strategy != boundary) { sp.x = side == 0 ? 1. : 0.; sp.y = yneighbour; sp.z = zneighbour; }
      } else {
        side = proy < 0. ? 0 : 1;
        findNeighbour(level, seq, 1, e.x, e.z, side, strategy);
        if (strategy != boundary) { sp.y = side == 0 ? 1. : 0.; sp.x = xneighbour; sp.z = zneighbour; }
      }
      face = dir * 2 + side;
      if (strategy != boundary) {  // checkPoint (equiSources.f90:2962)
        if (sp.x < 0. || sp.x > 1. || sp.y < 0. || sp.y > 1. || sp.z < 0. || sp.z > 1.) return ERR_CHECKPOINT;
      }
    } else if (radius * scale >= rmax[pixelLevel]) {
      strategy = split;
      length = 0.;
    } else {
      strategy = split;
      tmp = rmax[pixelLevel] - radius * scale;
      length = tmp;
      radius = radius + tmp / scale;
      sp = Pt{sp.x + tmp * prox, sp.y + tmp * proy, sp.z + tmp * proz};
    }
    return OK;
  }

  // absoluteCoordinates (equiSources.f90:3011-3047)
  void absoluteCoordinates(int level, const int* seq, Pt p) {
    for (int l = level; l > 0; l--) {
      p.x = seq[3 * l] == 1 ? 0.5 * p.x : 0.5 * p.x + 0.5;
      p.y = seq[3 * l + 1] == 1 ? 0.5 * p.y : 0.5 * p.y + 0.5;
      p.z = seq[3 * l + 2] == 1 ? 0.5 * p.z : 0.5 * p.z + 0.5;
    }
    xbase = ((double)(float)(seq[0] - 1) + p.x) / (double)(float)g.nx;
    ybase = ((double)(float)(seq[1] - 1) + p.y) / (double)(float)g.ny;
    zbase = ((double)(float)(seq[2] - 1) + p.z) / (double)(float)g.nz;
  }

  // localizeSplitContinuationCell (equiSources.f90:3049-3118)
  void localizeSplit(double x, double y, double z) {
    int i = (int)(x * g.nx) + 1, j = (int)(y * g.ny) + 1, k = (int)(z * g.nz) + 1;
    if (i > g.nx || j > g.ny || k > g.nz) { splitCell = -1; return; }  // coordinate exactly 1: the reference reads cell(nx+1)
    splitSeq[0] = i; splitSeq[1] = j; splitSeq[2] = k;
    int cell = g.base(i, j, k);
    double xn = x * (double)(float)g.nx - (double)(float)(i - 1);
    double yn = y * (double)(float)g.ny - (double)(float)(j - 1);
    double zn = z * (double)(float)g.nz - (double)(float)(k - 1);
    int level = 0;
    while (g.node[cell].refined()) {
      level++;
      i = xn < 0.5 ? 1 : 2; j = yn < 0.5 ? 1 : 2; k = zn < 0.5 ? 1 : 2;
      splitSeq[3 * level] = i; splitSeq[3 * level + 1] = j; splitSeq[3 * level + 2] = k;
      cell = g.kid(cell, i, j, k);
      xn = i == 1 ? 2. * xn : 2. * xn - 1.;
      yn = j == 1 ? 2. * yn : 2. * yn - 1.;
      zn = k == 1 ? 2. * zn : 2. * zn - 1.;
    }
    splitCell = cell;
    splitPoint = Pt{xn, yn, zn};
  }

  // startNewLongRay (equiSources.f90:3120-3385); pixel angles recomputed instead of cached (same values)
  int startNewLongRay(int startCell, Pt startPoint, int pixelLevel, int64_t irayStarting, int level, const int* startSeq,
                      double startRadius, double ndot1, double d1, double d2, double d3, double dD) {
    if (g.node[startCell].level != level) return ERR_ARG;
    double phi, theta;
    int st = pix2ang_nest(1 << (pixelLevel - 1), irayStarting - 1, phi, theta);
    if (st) return st;
    int cell = startCell;
    Pt cp = startPoint;
    double radius = startRadius, depth1 = d1, depth2 = d2, depth3 = d3, depthDust = dD;
    int newLevel = level;
    int seq[40];
    std::memcpy(seq, startSeq, sizeof(int) * (3 * level + 3));
    int strategy = proceed;
    const double boxOverNx = 0.;  (void)boxOverNx;
    while (strategy == proceed) {
      double oldRadius = radius, len;
      int face;
      st = drawSegment(cell, cp, phi, theta, pixelLevel, newLevel, seq, radius, strategy, len, face);
      if (st) return st;
      nseg++;
      if (trace && traceLen < traceCap)
        trace[traceLen++] = ((int64_t)g.node[cell].leaf << 32) | ((int64_t)pixelLevel << 28) | ((irayStarting - 1) << 8) | face;
      Zone& z = g.node[cell];
      double physicalCellSize = g.physicalBoxSize / ((double)(float)(1 << newLevel) * (double)(float)g.nx);
      double physicalLength = physicalCellSize * len;
      double tau1 = physicalLength * z.HI * (double)6.3e-18f;
      double tau2 = physicalLength * z.HeI * (double)7.42e-18f;
      double tau3 = physicalLength * z.HeII * (double)1.58e-18f;
      double tauDust;
      if (dustApproximation == 0) tauDust = 0.;
      else if (dustApproximation == 1) tauDust = physicalLength * z.HI * (double)5.4116737e-22f * z.abun2 / (double)0.2f;
      else tauDust = physicalLength * psi * z.rho / mh * (double)5.4116737e-22f * z.abun2 / (double)0.2f;
      for (int ir = 0; ir < 7; ir++) {
        double tmp = outputRadius[ir] * kpc;
        double t1 = oldRadius * g.physicalBoxSize / (double)(float)g.nx;
        double t2 = radius * g.physicalBoxSize / (double)(float)g.nx;
        if (tmp >= t1 && tmp <= t2) {
          double ratio = (tmp - t1) / (t2 - t1);
          ndotRemaining[ir] = ndotRemaining[ir] + ndot1 * xexp(-(ratio * (tau1 + tauDust) + depth1 + depthDust));
          if (ir == 6) {
            double o1 = ratio * tau1 + depth1, o2 = ratio * tau2 + depth2, o3 = ratio * tau3 + depth3;
            double oD = ratio * tauDust + depthDust;
            ndotDust = ndotDust + ndot1 * xexp(-oD);
            for (int ie = 0; ie < 300; ie++) {
              double a1 = T->outputSigma24[ie] / (double)6.30e-18f * o1;
              double a2 = T->outputSigma26[ie] / (double)7.42e-18f * o2;
              double a3 = T->outputSigma25[ie] / (double)1.58e-18f * o3;
              double aD = T->outputSigmaDust[ie] / (double)5.4116737e-22f * oD;
              ndotSpectrum[ie] = ndotSpectrum[ie] + ndot1 * xexp(-(a1 + a2 + a3 + aD));
            }
          }
        }
      }
      if (strategy == boundary) {
        double tmp = radius * g.physicalBoxSize / ((double)(float)g.nx * kpc);
        for (int ir = 0; ir < 7; ir++)
          if (outputRadius[ir] > tmp) ndotBoundary[ir] = ndotBoundary[ir] + ndot1;
      }
      if (std::fmin(std::fmin(depth1 + tau1, depth2 + tau2), std::fmin(depth3 + tau3, depthDust + tauDust)) > 100.) strategy = boundary;
      double a, b, ea, eb;
      double* rate = g.rate.data();
      const int64_t nl = (int64_t)g.leafNode.size(), lf = z.leaf;
      st = getRates(*T, dustApproximation, 1, depth1, depth2, depth3, depthDust, a, ea); if (st) return st;
      st = getRates(*T, dustApproximation, 1, depth1 + tau1, depth2, depth3, depthDust, b, eb); if (st) return st;
      rate[0 * nl + lf] = rate[0 * nl + lf] + ndot1 * (a - b);      // krate24
      rate[3 * nl + lf] = rate[3 * nl + lf] + ndot1 * (ea - eb);    // crate24
      st = getRates(*T, dustApproximation, 2, depth1, depth2, depth3, depthDust, a, ea); if (st) return st;
      st = getRates(*T, dustApproximation, 2, depth1, depth2 + tau2, depth3, depthDust, b, eb); if (st) return st;
      rate[2 * nl + lf] = rate[2 * nl + lf] + ndot1 * (a - b);      // krate26
      rate[5 * nl + lf] = rate[5 * nl + lf] + ndot1 * (ea - eb);    // crate26
      st = getRates(*T, dustApproximation, 3, depth1, depth2, depth3, depthDust, a, ea); if (st) return st;
      st = getRates(*T, dustApproximation, 3, depth1, depth2, depth3 + tau3, depthDust, b, eb); if (st) return st;
      rate[1 * nl + lf] = rate[1 * nl + lf] + ndot1 * (a - b);      // krate25
      rate[4 * nl + lf] = rate[4 * nl + lf] + ndot1 * (ea - eb);    // crate25
      depth1 = depth1 + tau1; depth2 = depth2 + tau2; depth3 = depth3 + tau3; depthDust = depthDust + tauDust;
      if (strategy == proceed) {
        cell = neighbourCell;
        newLevel = g.node[cell].level;
        std::memcpy(seq, neighbourSeq, sizeof(int) * (3 * newLevel + 3));
      }
    }
    if (strategy == split) {
      for (int iray = 1; iray <= 4; iray++) {
        const int childLevel = pixelLevel + 1;
        const int64_t childPix = 4 * irayStarting + iray - 4;  // 1-based
        double cphi, ctheta;
        st = pix2ang_nest(1 << (childLevel - 1), childPix - 1, cphi, ctheta);
        if (st) return st;
        if (childLevel > highestPixelLevel) highestPixelLevel = childLevel;
        const int count = 3 * g.node[cell].level + 3; (void)count;
        absoluteCoordinates(g.node[cell].level, seq, cp);
        xbase = xbase + radius / (double)(float)g.nx * (std::cos(cphi) * std::cos(ctheta) - std::cos(phi) * std::cos(theta));
        ybase = ybase + radius / (double)(float)g.ny * (std::sin(cphi) * std::cos(ctheta) - std::sin(phi) * std::cos(theta));
        zbase = zbase + radius / (double)(float)g.nz * (std::sin(ctheta) - std::sin(theta));
        if (xbase < 0. || xbase > 1. || ybase < 0. || ybase > 1. || zbase < 0. || zbase > 1.) {
          strategy = boundary;  // NOT reset for the following siblings (equiSources.f90:3336-3345)
          double tmp = radius * g.physicalBoxSize / ((double)(float)g.nx * kpc);
          for (int ir = 0; ir < 7; ir++)
            if (outputRadius[ir] > tmp) ndotBoundary[ir] = ndotBoundary[ir] + ndot1 / 4.;
        }
        if (strategy != boundary) {
          localizeSplit(xbase, ybase, zbase);
          if (splitCell < 0) return ERR_CHECKPOINT;
          if (splitPoint.x < 0. || splitPoint.x > 1. || splitPoint.y < 0. || splitPoint.y > 1. || splitPoint.z < 0. ||
              splitPoint.z > 1.) return ERR_CHECKPOINT;
          int sl = g.node[splitCell].level;
          int sseq[40];
          std::memcpy(sseq, splitSeq, sizeof(int) * (3 * sl + 3));
          st = startNewLongRay(splitCell, splitPoint, childLevel, childPix, sl, sseq, radius, ndot1 / 4., depth1, depth2,
                               depth3, depthDust);
          if (st) return st;
        }
      }
    }
    return OK;
  }
};

// highestPixelLevel (equiSources.f90:1266, :3316) of every source of the last pointSolve call on this thread
static thread_local std::vector<int32_t> g_lastHighestPixelLevel;
const std::vector<int32_t>& lastHighestPixelLevel() { return g_lastHighestPixelLevel; }

// equiSources.f90:1256-1370 for a list of sources given by host leaf and multiplicity
int pointSolve(Grid& g, const PointSpectra& S, int dustApproximation, int maxPixelLevel, int nsrc, const int32_t* srcLeaf,
               const int32_t* srcWeight, double* rates /* [6][nleaf] accumulated */, double* ndotRemaining /* [nsrc][7] */,
               double* ndotBoundary, double* ndotDust, double* ndotSpectrum /* [nsrc][300] */, int64_t* nsegOut,
               int64_t* trace, int64_t traceCap, int64_t* traceLen) {
  const int64_t nl = (int64_t)g.leafNode.size();
  g.rate.assign(rates, rates + 6 * nl);
  PointSolver ps(g);
  ps.dustApproximation = dustApproximation;
  ps.maxPixelLevel = maxPixelLevel;
  ps.trace = trace; ps.traceCap = traceCap;
  PointTables T;
  g_lastHighestPixelLevel.assign((size_t)std::max(nsrc, 0), 0);
  for (int s = 0; s < nsrc; s++) {
    if (srcWeight[s] <= 0) continue;
    for (int i = 0; i < 7; i++) ps.ndotRemaining[i] = ps.ndotBoundary[i] = 0.;
    ps.ndotDust = 0.;
    for (int i = 0; i < 300; i++) ps.ndotSpectrum[i] = 0.;
    ps.highestPixelLevel = 0;
    const int cell = g.leafNode[srcLeaf[s]];
    const Zone& host = g.node[cell];
    // metallicity bracket (equiSources.f90:1282-1293)
    double tmp = host.abun2 > (double)1.e-20f ? std::log10(host.abun2) : -20.;
    int iMetal = 1;
    while (tmp > S.metallicity[iMetal]) {  // metallicity(iMetal+1)
      iMetal++;
      if (iMetal + 1 == 5) break;
    }
    double coefMetal = (tmp - S.metallicity[iMetal - 1]) / (S.metallicity[iMetal] - S.metallicity[iMetal - 1]);
    coefMetal = std::fmin(std::fmax(0., coefMetal), 1.);
    stellarBetaTable(S, iMetal, coefMetal, T);
    ps.T = &T;
    // path of the host leaf (star%position)
    int seq[40], level = host.level, c = cell;
    for (int l = level; l >= 1; l--) {
      int p = g.node[c].parent, q = c - g.node[p].child;
      seq[3 * l] = (q >> 2) + 1; seq[3 * l + 1] = ((q >> 1) & 1) + 1; seq[3 * l + 2] = (q & 1) + 1;
      c = p;
    }
    seq[0] = c / (g.ny * g.nz) + 1; seq[1] = (c / g.nz) % g.ny + 1; seq[2] = c % g.nz + 1;
    const double ndot1 = (double)(float)srcWeight[s];
    for (int iray = 1; iray <= 12; iray++) {
      int st = ps.startNewLongRay(cell, Pt{0.5, 0.5, 0.5}, 1, iray, level, seq, 0., ndot1 / 12., 0., 0., 0., 0.);
      if (st) return st;
    }
    for (int i = 0; i < 7; i++) { ndotRemaining[s * 7 + i] = ps.ndotRemaining[i]; ndotBoundary[s * 7 + i] = ps.ndotBoundary[i]; }
    ndotDust[s] = ps.ndotDust;
    for (int i = 0; i < 300; i++) ndotSpectrum[(size_t)s * 300 + i] = ps.ndotSpectrum[i];
    g_lastHighestPixelLevel[s] = ps.highestPixelLevel;
  }
  std::memcpy(rates, g.rate.data(), sizeof(double) * 6 * nl);
  if (nsegOut) *nsegOut = ps.nseg;
  if (traceLen) *traceLen = ps.traceLen;
  return OK;
}

int pointTables(const PointSpectra& S, int iMetal, double coefMetal, double* out /* [6][11^4] R1..3, E1..3 */,
                double* totalIntegral, double* outputSigma /* [5][300] freq, s24, s25, s26, sDust */) {
  PointTables T;
  stellarBetaTable(S, iMetal, coefMetal, T);
  for (int r = 0; r < 3; r++) {
    std::memcpy(out + (size_t)r * nT, T.R[r].data(), sizeof(double) * nT);
    std::memcpy(out + (size_t)(3 + r) * nT, T.E[r].data(), sizeof(double) * nT);
  }
  if (totalIntegral) *totalIntegral = T.totalIntegral;
  if (outputSigma) {
    std::memcpy(outputSigma, T.outputFreq, 2400);
    std::memcpy(outputSigma + 300, T.outputSigma24, 2400);
    std::memcpy(outputSigma + 600, T.outputSigma25, 2400);
    std::memcpy(outputSigma + 900, T.outputSigma26, 2400);
    std::memcpy(outputSigma + 1200, T.outputSigmaDust, 2400);
  }
  return OK;
}

}  // namespace ftte
