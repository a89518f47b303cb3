// TEST INFRASTRUCTURE ONLY (see ftte_common.h).  extern "C" entry points of the CPU oracle, loaded with ctypes
// by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Never by the product.
#include "ftte_common.h"
#include "../radiativetransfer_b200/csrc/portable_math.h"

using namespace ftte;

extern "C" {

void* ftte_grid_create(int nx, int64_t nleaf, const int8_t* level, const double* HI, const double* HeI,
                       const double* HeII, const double* rho, const double* abun2, double boxSize, int* status) {
  Grid* g = new Grid();
  LeafInput in{level, HI, HeI, HeII, rho, abun2, nleaf};
  int st = buildGrid(*g, nx, boxSize, in);
  if (status) *status = st;
  if (st) { delete g; return nullptr; }
  return g;
}

void ftte_grid_destroy(void* h) { delete (Grid*)h; }

int64_t ftte_grid_nnodes(void* h) { return (int64_t)((Grid*)h)->node.size(); }

// update the absorber densities in place (outer transport<->chemistry iteration, equiSources.f90:3671-3673)
int ftte_grid_set_species(void* h, const double* HI, const double* HeI, const double* HeII) {
  Grid& g = *(Grid*)h;
  for (size_t l = 0; l < g.leafNode.size(); l++) {
    Zone& z = g.node[g.leafNode[l]];
    if (HI) z.HI = HI[l];
    if (HeI) z.HeI = HeI[l];
    if (HeII) z.HeII = HeII[l];
  }
  return OK;
}

// Diffuse solve over directions [rayBegin, rayEnd) (pass 0,-1 for all).  beta = [group][beta24, beta26, beta25].
// J1..J3: caller-allocated [nleaf], overwritten.  Optional trace of direction traceRay (-1 = none):
//   nbLeaf [3][nleaf] (xy, yz, xz upstream leaf; -1 boundary; -2 inactive), pattern [nx][12], izone, angles[2].
int ftte_diffuse(void* h, int nAngularLevel, const double* uvb, const double* beta, int64_t rayBegin,
                 int64_t rayEnd, double* J1, double* J2, double* J3, int64_t* nseg, int64_t traceRay,
                 int32_t* nbLeaf, double* patternOut, int32_t* izoneOut, double* anglesOut) {
  Grid& g = *(Grid*)h;
  DiffuseTrace tr{nbLeaf, patternOut, izoneOut, anglesOut};
  int st = diffuseSolve(g, nAngularLevel, uvb, beta, rayBegin, rayEnd, traceRay, traceRay >= 0 ? &tr : nullptr, nseg);
  for (size_t l = 0; l < g.leafNode.size(); l++) {
    const Zone& z = g.node[g.leafNode[l]];
    if (J1) J1[l] = z.Jmean[0];
    if (J2) J2[l] = z.Jmean[1];
    if (J3) J3[l] = z.Jmean[2];
  }
  return st;
}

// multi-threaded variant over an explicit list of HEALPix pixels; J = [3][nleaf] contiguous
int ftte_diffuse_mt(void* h, int nAngularLevel, const double* uvb, const double* beta, const int32_t* rays, int nrays,
                    int nthreads, double* J, int64_t* nseg) {
  return diffuseSolveThreaded(*(Grid*)h, nAngularLevel, uvb, beta, rays, nrays, nthreads, J, nseg);
}

// Point sources.  rates = [6][nleaf] (krate24, krate25, krate26, crate24, crate25, crate26), accumulated in place.
int ftte_point(void* h, int nWave, const double* wavelength, const double* lum, const double* metallicity,
               double coefSpectrum, const double* aDust, int dustApproximation, int maxPixelLevel, int nsrc,
               const int32_t* srcLeaf, const int32_t* srcWeight, double* rates, double* ndotRemaining,
               double* ndotBoundary, double* ndotDust, double* ndotSpectrum, int64_t* nseg, int64_t* trace,
               int64_t traceCap, int64_t* traceLen) {
  PointSpectra S{nWave, wavelength, lum, metallicity, coefSpectrum, aDust};
  return pointSolve(*(Grid*)h, S, dustApproximation, maxPixelLevel, nsrc, srcLeaf, srcWeight, rates, ndotRemaining,
                    ndotBoundary, ndotDust, ndotSpectrum, nseg, trace, traceCap, traceLen);
}

// highestPixelLevel of the sources of the calling thread's last ftte_point call (equiSources.f90:1353-1357, 'src:' line)
int ftte_point_highest_pixel_level(int32_t* out, int n) {
  const auto& v = ftte::lastHighestPixelLevel();
  for (int i = 0; i < n; i++) out[i] = i < (int)v.size() ? v[i] : 0;
  return (int)v.size();
}

// host build of the portable exp/log (checked against a 50-digit evaluation in tests/test_portable_math.py)
void ftte_pm_eval(int64_t n, const double* x, double* expOut, double* logOut) {
  for (int64_t i = 0; i < n; i++) {
    if (expOut) expOut[i] = rtb_pm::pm_exp(x[i]);
    if (logOut) logOut[i] = rtb_pm::pm_log(x[i]);
  }
}

void ftte_set_portable_math(int on) { setPortableMath(on); }

// solveRateEquations on flattened leaf arrays (HI, HeI, HeII are updated in place)
int ftte_chemistry(int64_t nleaf, int nx, double box, const int8_t* level, const double* rho, const double* tgas, double* HI,
                   double* HeI, double* HeII, const double* rates, const double* J, const double* ksi,
                   const double* uniform, int nratec, double logtem0, double logtem9, double dlogtem, const double* k,
                   double* maxChange) {
  return chemistrySolve(nleaf, nx, box, level, rho, tgas, HI, HeI, HeII, rates, J, ksi, uniform, nratec, logtem0, logtem9,
                        dlogtem, k, k + nratec, k + 2 * (size_t)nratec, k + 3 * (size_t)nratec, k + 4 * (size_t)nratec,
                        k + 5 * (size_t)nratec, maxChange);
}

// computeMass (equiSources.f90:4369-4393): serial accumulation in leaf order, the order the recursion visits leaves
void ftte_compute_mass(int64_t nleaf, int nx, double box, const int8_t* level, const double* HI, const double* rho,
                       double* out) {
  double neutralHydrogenMass = 0., totalHydrogenMass = 0.;
  for (int64_t c = 0; c < nleaf; c++) {
    const double size = box / ((double)(float)(1 << level[c]) * (double)(float)nx);   // :4389, float() = real*4
    const double physicalCellSizeCube = size * size * size;                           // x**3
    neutralHydrogenMass = neutralHydrogenMass + HI[c] * mh * physicalCellSizeCube / msun;          // :4390
    totalHydrogenMass = totalHydrogenMass + psi * rho[c] * physicalCellSizeCube / msun;            // :4391
  }
  out[0] = neutralHydrogenMass;
  out[1] = totalHydrogenMass;
}

int ftte_point_tables(int nWave, const double* wavelength, const double* lum, const double* metallicity,
                      double coefSpectrum, const double* aDust, int iMetal, double coefMetal, double* out,
                      double* totalIntegral, double* outputSigma) {
  PointSpectra S{nWave, wavelength, lum, metallicity, coefSpectrum, aDust};
  return pointTables(S, iMetal, coefMetal, out, totalIntegral, outputSigma);
}

int ftte_direction(int nAngularLevel, int64_t iray, int32_t* izone, double* phi, double* theta) {
  int iz = 0;
  int st = directionSetup(nAngularLevel, iray, iz, *phi, *theta);
  *izone = iz;
  return st;
}

int ftte_pix2ang_nest(int nside, int64_t ipix, double* phi, double* theta) { return pix2ang_nest(nside, ipix, *phi, *theta); }

void ftte_rotate_indices(int i, int j, int k, int nx, int ny, int nz, int izone, int32_t* out) {
  int a, b, c;
  rotateIndices(i, j, k, nx, ny, nz, izone, a, b, c);
  out[0] = a; out[1] = b; out[2] = c;
}

}  // extern "C"
