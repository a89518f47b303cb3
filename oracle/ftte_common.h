// TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the razoumov/radiativeTransfer hot path.
// Nothing under oracle/ may be imported, linked or executed by the product (radiativetransfer_b200/);
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures, and cannot be compiled in
// this image (no Fortran compiler, no HDF4).  This restatement follows the cited reference lines one by
// one and is pinned only by the analytic known-answer tests in tests/test_oracle_*.py.
//
// Shared constants and grid (octree) structures.  All file:line citations are relative to /root/reference.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

namespace ftte {

// ---- constants: Fortran real literals without a d-exponent are single precision, then widened
//      (definitionsModule.f90:8-41, 261) -------------------------------------------------------------
static const double pi = (double)3.141592654f;                 // definitionsModule.f90:8
static const double halfPi = 0.5 * pi;                         // :9
static const double twoPi = 2.0 * pi;                          // :10
static const double hp = (double)6.6260693e-27f;               // :15
static const double clight = (double)2.99792458e10f;           // :17
static const double yr = 31557600.0;                           // :18 (integer literal)
static const double Myr = (double)1.e6f * yr;                  // :20
static const double pc = (double)3.08568025e18f;               // :21
static const double kpc = (double)1.e3f * pc;                  // :22
static const double angstrom = (double)1.e-8f;                 // :24
static const double mp = (double)1.6726231e-24f;               // :25
static const double mn = (double)1.67492728e-24f;              // :26
static const double mh = mp;                                   // :27
static const double mhe = 2.0 * (mp + mn);                     // :28
static const double msun = (double)1.98892e33f;                // :29
static const double hydrogenIonization = (double)13.598f;      // :30
static const double singleHeliumIonization = (double)24.587f;  // :31
static const double doubleHeliumIonization = (double)54.418f;  // :32
static const double nu1 = hydrogenIonization, nu2 = singleHeliumIonization, nu3 = doubleHeliumIonization;
static const double eV_to_erg = 1.60217646e-12;                // :36 (true double)
static const double eV_to_Hz = eV_to_erg / hp;                 // :38
static const double psi = (double)0.76f;                       // :261

enum { xyEnd = 1, yzEnd = 2, xzEnd = 3 };                      // definitionsModule.f90:158
enum { proceed = 1, split = 2, boundary = 3 };                 // :232

// status codes returned instead of the reference's `write; stop`
enum Status {
  OK = 0,
  ERR_PHI = 1,            // equiSources.f90:1413 'error in phi'
  ERR_THETA = 2,          // :1426 'error in theta'
  ERR_THETA_OR_PHI = 3,   // :1449 'error in theta or phi'
  ERR_PATTERN_RANGE = 4,  // transportRoutinesModule.f90:33,60,183; equiSources.f90:1523
  ERR_TOP_SELECTOR = 5,   // 'error in xyTop/xzTop/yzTop' (selector 0 without a coarser neighbour)
  ERR_RAY_INACTIVE = 6,   // 'Error: xzRay should be active'
  ERR_INTENSITY_GUARD = 7,// transportRoutinesModule.f90:680-688 (|sum Iout| >= 1e-20 in a refined leaf)
  ERR_ANGLE_LARGE = 8,    // equiSources.f90:2224
  ERR_LEVELS = 9,         // readCellArray.f90:181 'error in levels'
  ERR_CHECKPOINT = 10,    // equiSources.f90:2962 checkPoint
  ERR_IDEPTH = 11,        // equiSources.f90:4196
  ERR_ARG = 12
};

// ---- octree node: the subset of zoneType (definitionsModule.f90:163-180) the hot path touches ----------
struct Zone {
  double Iout[3][3];      // rt%{xy,yz,xz}Ray%Iout{1,2,3}; first index: 0=xy 1=yz 2=xz  (ray id - 1)
  double rho, HI, HeI, HeII, abun2;
  double kappa[3];
  double Jmean[3];
  int32_t parent;         // node index, -1 = baseGrid
  int32_t child;          // index of first of 8 children (i,j,k order, i slowest), -1 = leaf
  int32_t nb[3];          // xyNeighbour, yzNeighbour, xzNeighbour (indexed by ray id - 1), leaf NODE index
  int32_t pattern;        // index into the per-direction pattern pool
  int32_t leaf;           // leaf number in writeCell pre-order (equiSources.f90:4044-4079), -1 if refined
  bool nbPresent[3];
  int8_t level;
  bool refined() const { return child >= 0; }
};

struct Grid {
  int nx = 0, ny = 0, nz = 0;
  double physicalBoxSize = 0;
  std::vector<Zone> node;           // base cells first: index ((i-1)*ny + (j-1))*nz + (k-1)
  std::vector<int32_t> leafNode;    // leaf number -> node index
  std::vector<double> rate;         // [6][nleaf] krate24, krate25, krate26, crate24, crate25, crate26 (point sources)
  std::vector<Grid> threadCopy;     // private octree copies of the worker threads (diffuseSolveThreaded), made once
  int maxLevel = 0;
  int base(int i, int j, int k) const { return ((i - 1) * ny + (j - 1)) * nz + (k - 1); }
  // child (i,j,k) in 1..2 of a refined node
  int kid(int n, int i, int j, int k) const { return node[n].child + (i - 1) * 4 + (j - 1) * 2 + (k - 1); }
};

// flattened leaf arrays in writeCell pre-order (the C-ABI boundary of the product uses the same layout)
struct LeafInput {
  const int8_t* level;
  const double *HI, *HeI, *HeII, *rho, *abun2;
  int64_t nleaf;
};

struct DiffuseTrace {      // optional per-direction exports for the bit-exact traversal checks
  int32_t* nbLeaf;         // [3][nleaf] upstream LEAF numbers (xy, yz, xz), -1 = boundary, -2 = ray inactive
  double* patternOut;      // [nx][12] base-layer patterns: xy(x0,y0,len) xz(x0,z0,len) yz(y0,z0,len) xyTop xzTop yzTop
  int32_t* izoneOut;
  double* anglesOut;       // phi, theta (local)
};

// stellar population synthesis inputs of the point-source path (equiSources.f90:840-892, dustModule.f90:15-24):
// none of these data files ship with the reference, so they are arguments
struct PointSpectra {
  int nWave;                  // nWavelengths (1221 in the reference)
  const double* wavelength;   // [nWave] cm, increasing
  const double* lum;          // [5 metallicities][2 time slices: iSpectrum, iSpectrum+1][nWave] log10(erg/s/A)
  const double* metallicity;  // [5] log10 Z
  double coefSpectrum;        // time interpolation weight (equiSources.f90:1241-1242)
  const double* aDust;        // [7][5] SMC extinction-fit parameters a_smc(i, 1..5)
};
int pointSolve(Grid& g, const PointSpectra& S, int dustApproximation, int maxPixelLevel, int nsrc, const int32_t* srcLeaf,
               const int32_t* srcWeight, double* rates, double* ndotRemaining, double* ndotBoundary, double* ndotDust,
               double* ndotSpectrum, int64_t* nsegOut, int64_t* trace, int64_t traceCap, int64_t* traceLen);
void setPortableMath(int on);
const std::vector<int32_t>& lastHighestPixelLevel();   // per source of the calling thread's last pointSolve
int chemistrySolve(int64_t nleaf, int nx, double physicalBoxSize, const int8_t* level, const double* rho,
                   const double* tgas, double* HI, double* HeI, double* HeII, const double* rates, const double* J,
                   const double* ksi, const double* uniform, int nratec, double logtem0, double logtem9, double dlogtem,
                   const double* k1a, const double* k2a, const double* k3a, const double* k4a, const double* k5a,
                   const double* k6a, double* maxChange);
int pointTables(const PointSpectra& S, int iMetal, double coefMetal, double* out, double* totalIntegral,
                double* outputSigma);

int buildGrid(Grid& g, int nx, double boxSize, const LeafInput& in);
int diffuseSolve(Grid& g, int nAngularLevel, const double* uvb, const double* beta, int64_t rayBegin, int64_t rayEnd,
                 int64_t traceRay, DiffuseTrace* tr, int64_t* nsegOut);
// several host threads, each sweeping its own share of `rays` on a private copy of the octree (the reference itself
// is serial; this is the "all host threads" variant used by bench.py --impl reference)
int diffuseSolveThreaded(Grid& g, int nAngularLevel, const double* uvb, const double* beta, const int32_t* rays,
                         int nrays, int nthreads, double* J /* [3][nleaf] */, int64_t* nsegOut);
int directionSetup(int nAngularLevel, int64_t iray, int& izone, double& phi, double& theta);
int pix2ang_nest(int nside, int64_t ipix, double& phi, double& theta);
void rotateIndices(int i, int j, int k, int nx, int ny, int nz, int izone, int& ic, int& jc, int& kc);

}  // namespace ftte
