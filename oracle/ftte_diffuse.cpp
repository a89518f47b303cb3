// TEST INFRASTRUCTURE ONLY (see ftte_common.h).  PARITY UNPINNED (no reference golden vectors exist).
//
// CPU restatement of the diffuse (UV-background) sweep of razoumov/radiativeTransfer:
//   equiSources.f90:1372-1808 (driver block), :2118-2361 (HEALPix + angle utilities), :4956-4983 (opacities)
//   transportRoutinesModule.f90:7-85 (setPattern), :121-218 (setRaysRefined), :264-558 (neighbours),
//   :560-963 (transport), :1036-1054 (computeCellIntensity);  rotateIndicesModule.f90:7-113.
// Single thread, libm exp/log/trig (what gfortran calls), no FMA contraction (build with -ffp-contract=off).
#include "ftte_common.h"

#include <thread>

namespace ftte {

// ------------------------------------------------------------------------------------------------------
// grid construction from the flattened leaf array (leaf pre-order: equiSources.f90:4830-4836, :4044-4079;
// tree rebuilt from `level` alone exactly as readCellArray.f90:154-187 createFullyThreadedStructure)
// ------------------------------------------------------------------------------------------------------
static int createFullyThreaded(Grid& g, int n, int level, int64_t& icosmic, const LeafInput& in) {
  if (icosmic >= in.nleaf) return ERR_LEVELS;
  int lv = in.level[icosmic];
  if (lv > g.maxLevel) g.maxLevel = lv;
  if (lv == level) {
    Zone& z = g.node[n];
    z.leaf = (int32_t)icosmic;
    z.HI = in.HI[icosmic];
    z.HeI = in.HeI ? in.HeI[icosmic] : 0.0;
    z.HeII = in.HeII ? in.HeII[icosmic] : 0.0;
    z.rho = in.rho ? in.rho[icosmic] : 0.0;
    z.abun2 = in.abun2 ? in.abun2[icosmic] : 0.0;
    g.leafNode[icosmic] = n;
    icosmic++;
    return OK;
  }
  if (lv < level) return ERR_LEVELS;
  int c = (int)g.node.size();
  g.node[n].child = c;
  g.node.resize(g.node.size() + 8);
  for (int q = 0; q < 8; q++) {
    Zone& z = g.node[c + q];
    std::memset(&z, 0, sizeof(Zone));
    z.parent = n;
    z.child = -1;
    z.leaf = -1;
    z.level = (int8_t)(level + 1);
    z.pattern = -1;
  }
  for (int q = 0; q < 8; q++) {
    int st = createFullyThreaded(g, c + q, level + 1, icosmic, in);
    if (st) return st;
  }
  return OK;
}

int buildGrid(Grid& g, int nx, double boxSize, const LeafInput& in) {
  g.nx = g.ny = g.nz = nx;
  g.physicalBoxSize = boxSize;
  g.maxLevel = 0;
  g.node.clear();
  g.node.reserve((size_t)(in.nleaf + in.nleaf / 7 + 16));
  g.node.resize((size_t)nx * nx * nx);
  g.leafNode.assign((size_t)in.nleaf, -1);
  for (auto& z : g.node) {
    std::memset(&z, 0, sizeof(Zone));
    z.parent = -1;
    z.child = -1;
    z.leaf = -1;
    z.level = 0;
    z.pattern = -1;
  }
  int64_t icosmic = 0;
  for (int i = 1; i <= nx; i++)
    for (int j = 1; j <= nx; j++)
      for (int k = 1; k <= nx; k++) {
        int st = createFullyThreaded(g, g.base(i, j, k), 0, icosmic, in);
        if (st) return st;
      }
  if (icosmic != in.nleaf) return ERR_LEVELS;
  return OK;
}

// ------------------------------------------------------------------------------------------------------
// rotateIndicesModule.f90:7-113
// ------------------------------------------------------------------------------------------------------
void rotateIndices(int i, int j, int k, int nx, int ny, int nz, int izone, int& ic, int& jc, int& kc) {
  switch (izone) {
    case 1: ic = i; jc = j; kc = k; break;
    case 2: ic = j; jc = k; kc = i; break;
    case 3: ic = k; jc = i; kc = j; break;
    case 4: ic = i; jc = k; kc = nz + 1 - j; break;
    case 5: ic = j; jc = i; kc = nz + 1 - k; break;
    case 6: ic = k; jc = j; kc = nz + 1 - i; break;
    case 7: ic = i; jc = ny + 1 - j; kc = nz + 1 - k; break;
    case 8: ic = j; jc = ny + 1 - k; kc = nz + 1 - i; break;
    case 9: ic = k; jc = ny + 1 - i; kc = nz + 1 - j; break;
    case 10: ic = i; jc = ny + 1 - k; kc = j; break;
    case 11: ic = j; jc = ny + 1 - i; kc = k; break;
    case 12: ic = k; jc = ny + 1 - j; kc = i; break;
    case 13: ic = nx + 1 - i; jc = j; kc = k; break;
    case 14: ic = nx + 1 - j; jc = k; kc = i; break;
    case 15: ic = nx + 1 - k; jc = i; kc = j; break;
    case 16: ic = nx + 1 - i; jc = k; kc = nz + 1 - j; break;
    case 17: ic = nx + 1 - j; jc = i; kc = nz + 1 - k; break;
    case 18: ic = nx + 1 - k; jc = j; kc = nz + 1 - i; break;
    case 19: ic = nx + 1 - i; jc = ny + 1 - j; kc = nz + 1 - k; break;
    case 20: ic = nx + 1 - j; jc = ny + 1 - k; kc = nz + 1 - i; break;
    case 21: ic = nx + 1 - k; jc = ny + 1 - i; kc = nz + 1 - j; break;
    case 22: ic = nx + 1 - i; jc = ny + 1 - k; kc = j; break;
    case 23: ic = nx + 1 - j; jc = ny + 1 - i; kc = k; break;
    case 24: ic = nx + 1 - k; jc = ny + 1 - j; kc = i; break;
    default: ic = jc = kc = 0;
  }
}

// ------------------------------------------------------------------------------------------------------
// HEALPix nested pixel centre + fixed rotations (equiSources.f90:2118-2361)
// ------------------------------------------------------------------------------------------------------
static int pix2x[1024], pix2y[1024];
static bool pix2xyReady = false;

static void mk_pix2xy() {  // equiSources.f90:2233-2275
  for (int kpix = 0; kpix <= 1023; kpix++) {
    int jpix = kpix, ix = 0, iy = 0, ip = 1;
    while (jpix != 0) {
      int id = jpix % 2; jpix /= 2; ix = id * ip + ix;
      id = jpix % 2; jpix /= 2; iy = id * ip + iy;
      ip = 2 * ip;
    }
    pix2x[kpix] = ix;
    pix2y[kpix] = iy;
  }
  pix2xyReady = true;
}

static double arcsin_(double x) {  // equiSources.f90:2277-2295
  if (x > 1.0) return halfPi;
  if (x < -1.0) return -halfPi;
  return std::asin(x);
}

static double getAngle(double cosphi, double sinphi) {  // equiSources.f90:2337-2361
  double phi = arcsin_(sinphi);
  if (cosphi > 0.) {
    if (sinphi > 0.) phi = phi; else phi = twoPi + phi;
  } else {
    phi = pi - phi;
  }
  return phi;
}

static void rotateAngles(double& phi, double& theta) {  // equiSources.f90:2297-2335
  double phi0 = phi, theta0 = theta;
  double rotationAngle = (double)0.111f;
  theta = arcsin_(std::cos(theta0) * std::sin(phi0) * std::sin(rotationAngle) + std::sin(theta0) * std::cos(rotationAngle));
  double cosphi = std::cos(theta0) * std::cos(phi0) / std::cos(theta);
  double sinphi = (std::cos(theta0) * std::sin(phi0) * std::cos(rotationAngle) - std::sin(theta0) * std::sin(rotationAngle)) / std::cos(theta);
  phi = getAngle(cosphi, sinphi);
  phi0 = phi; theta0 = theta;
  rotationAngle = (double)0.222f;
  theta = arcsin_(std::cos(theta0) * std::cos(phi0) * std::sin(rotationAngle) + std::sin(theta0) * std::cos(rotationAngle));
  cosphi = (std::cos(theta0) * std::cos(phi0) * std::cos(rotationAngle) - std::sin(theta0) * std::sin(rotationAngle)) / std::cos(theta);
  sinphi = std::cos(theta0) * std::sin(phi0) / std::cos(theta);
  phi = getAngle(cosphi, sinphi);
}

int pix2ang_nest(int nside, int64_t ipix, double& phi, double& theta) {  // equiSources.f90:2118-2231
  static const int jrll[12] = {2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4};
  static const int jpll[12] = {1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7};
  int64_t nsideLong = nside;
  if (nside < 1 || nside > 8192 * 4) return ERR_ARG;
  int64_t npix = 12 * nsideLong * nsideLong;
  if (ipix < 0 || ipix > npix - 1) return ERR_ARG;
  if (!pix2xyReady) mk_pix2xy();
  double fn = (double)(float)nside;
  double fact1 = 1.0 / (3.0 * fn * fn);
  double fact2 = 2.0 / (3.0 * fn);
  int nl4 = 4 * nside;
  int64_t npface = nsideLong * nsideLong;
  int face_num = (int)(ipix / npface);
  int64_t ipf = ipix % npface;
  int64_t ip_low = ipf % 1024;
  int64_t ip_trunc = ipf / 1024;
  int64_t ip_med = ip_trunc % 1024;
  int ip_hi = (int)(ip_trunc / 1024);
  int ix = 1024 * pix2x[ip_hi] + 32 * pix2x[ip_med] + pix2x[ip_low];
  int iy = 1024 * pix2y[ip_hi] + 32 * pix2y[ip_med] + pix2y[ip_low];
  int jrt = ix + iy;
  int jpt = ix - iy;
  int jr = jrll[face_num] * nside - jrt - 1;
  int nr = nside;
  double z = (double)(float)(2 * nside - jr) * fact2;
  int kshift = (jr - nside) % 2;  // Fortran MOD == C % (sign of dividend)
  if (jr < nside) {
    nr = jr;
    z = 1.0 - (double)((float)nr * (float)nr) * fact1;
    kshift = 0;
  } else if (jr > 3 * nside) {
    nr = nl4 - jr;
    z = -1.0 + (double)((float)nr * (float)nr) * fact1;
    kshift = 0;
  }
  theta = std::acos(z) - halfPi;
  int jp = (jpll[face_num] * nr + jpt + 1 + kshift) / 2;  // truncation toward zero, as Fortran
  if (jp > nl4) jp = jp - nl4;
  if (jp < 1) jp = jp + nl4;
  phi = (double)((float)jp - (float)(kshift + 1) * 0.5f) * halfPi / (double)(float)nr;
  while (phi > twoPi) phi = phi - twoPi;
  while (phi < 0.) phi = phi + twoPi;
  rotateAngles(phi, theta);
  if (phi > 2.0 * pi) return ERR_ANGLE_LARGE;
  return OK;
}

// ------------------------------------------------------------------------------------------------------
// patterns (definitionsModule.f90:141-152)
// ------------------------------------------------------------------------------------------------------
struct Pattern {
  double xy_x0, xy_y0, xy_len;
  double xz_x0, xz_z0, xz_len;
  double yz_y0, yz_z0, yz_len;
  bool xzRayActive, yzRayActive, refined;
  int8_t xyTop, xzTop, yzTop;
  // setRaysRefined allocates cell(2,2,2) but all four (j,k) members of a sub-layer are copies of one
  // pattern (transportRoutinesModule.f90:191-196): only rotated i = 1 (lower) / 2 (upper) matters.
  int32_t sub[2];
};

// transportRoutinesModule.f90:7-85
static int setPattern(Pattern& p, double phi, double theta) {
  double tmp1 = 1. / std::sin(theta);
  double tmp2 = (1. - p.xy_x0) / (std::cos(phi) * std::cos(theta));
  double tmp3 = (1. - p.xy_y0) / (std::sin(phi) * std::cos(theta));
  if (tmp1 < std::fmin(tmp2, tmp3)) {
    p.xy_len = tmp1;
    p.xzRayActive = false; p.yzRayActive = false;
    p.xyTop = xyEnd; p.xzTop = 0; p.yzTop = 0;
  } else if (tmp2 < std::fmin(tmp1, tmp3)) {
    p.xy_len = tmp2;
    p.yzRayActive = true;
    p.yz_y0 = (1. - p.xy_x0) * std::tan(phi) + p.xy_y0;
    p.yz_z0 = p.xy_len * std::sin(theta);
    if (p.yz_y0 > 1. || p.yz_z0 > 1.) return ERR_PATTERN_RANGE;
    double tmpa1 = (1. - p.yz_z0) / std::sin(theta);
    double tmpa2 = (1. - p.yz_y0) / (std::sin(phi) * std::cos(theta));
    if (tmpa1 < tmpa2) {
      p.yz_len = tmpa1;
      p.xzRayActive = false;
      p.xyTop = yzEnd; p.xzTop = 0; p.yzTop = xyEnd;
    } else {
      p.yz_len = tmpa2;
      p.xzRayActive = true;
      p.xz_x0 = (1. - p.yz_y0) / std::tan(phi);
      p.xz_z0 = p.yz_z0 + tmpa2 * std::sin(theta);
      p.xz_len = (1. - p.xz_z0) / std::sin(theta);
      p.xyTop = xzEnd; p.xzTop = yzEnd; p.yzTop = xyEnd;
    }
  } else {
    p.xy_len = tmp3;
    p.xzRayActive = true;
    p.xz_x0 = (1. - p.xy_y0) / std::tan(phi) + p.xy_x0;
    p.xz_z0 = tmp3 * std::sin(theta);
    if (p.xz_x0 > 1. || p.xz_z0 > 1.) return ERR_PATTERN_RANGE;
    double tmpb1 = (1. - p.xz_z0) / std::sin(theta);
    double tmpb2 = (1. - p.xz_x0) / (std::cos(phi) * std::cos(theta));
    if (tmpb1 < tmpb2) {
      p.xz_len = tmpb1;
      p.yzRayActive = false;
      p.xyTop = xzEnd; p.xzTop = xyEnd; p.yzTop = 0;
    } else {
      p.xz_len = tmpb2;
      p.yzRayActive = true;
      p.yz_y0 = (1. - p.xz_x0) * std::tan(phi);
      p.yz_z0 = p.xz_len * std::sin(theta) + p.xz_z0;
      p.yz_len = (1. - p.yz_z0) / std::sin(theta);
      p.xyTop = yzEnd; p.xzTop = xyEnd; p.yzTop = xzEnd;
    }
  }
  return OK;
}

// entry point of the layer above `below` (equiSources.f90:1507-1522 == transportRoutinesModule.f90:167-182)
static int continueAbove(const Pattern& below, Pattern& cur, double phi, double theta) {
  switch (below.xyTop) {
    case xyEnd:
      cur.xy_x0 = below.xy_x0 + std::cos(phi) / std::tan(theta);
      cur.xy_y0 = below.xy_y0 + std::sin(phi) / std::tan(theta);
      break;
    case xzEnd:
      cur.xy_x0 = below.xz_x0 + below.xz_len * std::cos(theta) * std::cos(phi);
      cur.xy_y0 = below.xz_len * std::cos(theta) * std::sin(phi);
      break;
    case yzEnd:
      cur.xy_x0 = below.yz_len * std::cos(theta) * std::cos(phi);
      cur.xy_y0 = below.yz_y0 + below.yz_len * std::cos(theta) * std::sin(phi);
      break;
    default:
      return ERR_TOP_SELECTOR;
  }
  if (cur.xy_x0 > 1. || cur.xy_y0 > 1.) return ERR_PATTERN_RANGE;
  return OK;
}

struct Sweep {
  Grid* g;
  std::vector<Pattern> pool;  // per-direction pattern storage (pattern(:) array + lazily built sub-trees)
  int is[3][3][3], js[3][3][3], ks[3][3][3];  // 1-based (2,2,2) tables, equiSources.f90:1485-1491
  int izone;
  double phi, theta;
  double uvb[3];
  int64_t nseg;               // segment updates performed (metric unit)
  int status;
};

static int newPattern(Sweep& s) {
  Pattern p;
  std::memset(&p, 0, sizeof(p));
  p.sub[0] = p.sub[1] = -1;
  s.pool.push_back(p);
  return (int)s.pool.size() - 1;
}

// transportRoutinesModule.f90:121-218
static int setRaysRefined(Sweep& s, int parentCell, int parentPattern) {
  Grid& g = *s.g;
  if (!s.pool[parentPattern].refined) {
    int lo = newPattern(s), up = newPattern(s);
    Pattern& pp = s.pool[parentPattern];
    pp.sub[0] = lo; pp.sub[1] = up;
    pp.refined = true;
    Pattern& cl = s.pool[lo];
    if (pp.xy_x0 < 0.5) cl.xy_x0 = 2. * pp.xy_x0; else cl.xy_x0 = 2. * pp.xy_x0 - 1.;
    if (pp.xy_y0 < 0.5) cl.xy_y0 = 2. * pp.xy_y0; else cl.xy_y0 = 2. * pp.xy_y0 - 1.;
    int st = setPattern(cl, s.phi, s.theta);
    if (st) return st;
    Pattern& cu = s.pool[up];
    st = continueAbove(cl, cu, s.phi, s.theta);
    if (st) return st;
    st = setPattern(cu, s.phi, s.theta);
    if (st) return st;
  }
  for (int i = 1; i <= 2; i++)
    for (int j = 1; j <= 2; j++)
      for (int k = 1; k <= 2; k++) {
        int cp = s.pool[parentPattern].sub[i - 1];
        int cc = g.kid(parentCell, s.is[i][j][k], s.js[i][j][k], s.ks[i][j][k]);
        g.node[cc].pattern = cp;
        if (g.node[cc].refined()) {
          int st = setRaysRefined(s, cc, cp);
          if (st) return st;
        }
      }
  return OK;
}

// transportRoutinesModule.f90:455-558 get{XY,XZ,YZ}Neighbour: descend the container to a leaf.
// `ray`: 0 = xy (face point (x0,y0) -> children (2, y, x)), 2 = xz ((x0,z0) -> (z, 2, x)), 1 = yz ((y0,z0) -> (z, y, 2)).
static void getNeighbour(Sweep& s, int cell, int ray, int container, double a, double b) {
  Grid& g = *s.g;
  while (g.node[container].refined()) {
    int ia = (a <= 0.5) ? 1 : 2, ib = (b <= 0.5) ? 1 : 2;
    int i, j, k;
    if (ray == 0) { i = 2; j = ib; k = ia; }        // a = x0, b = y0
    else if (ray == 2) { i = ib; j = 2; k = ia; }   // a = x0, b = z0
    else { i = ib; j = ia; k = 2; }                 // a = y0, b = z0
    container = g.kid(container, s.is[i][j][k], s.js[i][j][k], s.ks[i][j][k]);
    a = (ia == 1) ? 2. * a : 2. * a - 1.;
    b = (ib == 1) ? 2. * b : 2. * b - 1.;
  }
  g.node[cell].nbPresent[ray] = true;
  g.node[cell].nb[ray] = container;
}

// transportRoutinesModule.f90:264-418
static void findNeighbours(Sweep& s, int cell, int level, const int* callSequence) {
  Grid& g = *s.g;
  const Pattern& p = s.pool[g.node[cell].pattern];
  // ray 0: xy (walk while i == 1), ray 2: xz (while j == 1), ray 1: yz (while k == 1)
  for (int ray = 0; ray < 3; ray++) {
    double a, b;
    if (ray == 0) { a = p.xy_x0; b = p.xy_y0; }
    else if (ray == 2) { if (!p.xzRayActive) continue; a = p.xz_x0; b = p.xz_z0; }
    else { if (!p.yzRayActive) continue; a = p.yz_y0; b = p.yz_z0; }
    int coarser = cell;
    bool found = false;
    for (int ilevel = level; ilevel >= 0; ilevel--) {
      int i = callSequence[3 * ilevel], j = callSequence[3 * ilevel + 1], k = callSequence[3 * ilevel + 2];
      coarser = g.node[coarser].parent;
      int lead = (ray == 0) ? i : (ray == 2 ? j : k);
      if (lead > 1) {
        int i2 = i - (ray == 0), j2 = j - (ray == 2), k2 = k - (ray == 1);
        int container;
        if (ilevel == 0) {
          int ic, jc, kc;
          rotateIndices(i2, j2, k2, g.nx, g.ny, g.nz, s.izone, ic, jc, kc);
          container = g.base(ic, jc, kc);
        } else {
          container = g.kid(coarser, s.is[i2][j2][k2], s.js[i2][j2][k2], s.ks[i2][j2][k2]);
        }
        getNeighbour(s, cell, ray, container, a, b);
        found = true;
        break;
      }
      // rescale the face point to the parent's units
      if (ray == 0) {        // (x,y): y follows j, x follows k
        b = (j == 1) ? b / 2. : b / 2. + 0.5;
        a = (k == 1) ? a / 2. : a / 2. + 0.5;
      } else if (ray == 2) { // (x,z): z follows i, x follows k
        b = (i == 1) ? b / 2. : b / 2. + 0.5;
        a = (k == 1) ? a / 2. : a / 2. + 0.5;
      } else {               // (y,z): z follows i, y follows j
        b = (i == 1) ? b / 2. : b / 2. + 0.5;
        a = (j == 1) ? a / 2. : a / 2. + 0.5;
      }
    }
    if (!found) g.node[cell].nbPresent[ray] = false;
  }
}

// transportRoutinesModule.f90:421-453
static void localizeCellFindNeighbours(Sweep& s, int cell, int level, int* callSequence) {
  Grid& g = *s.g;
  if (g.node[cell].refined()) {
    for (int i = 1; i <= 2; i++)
      for (int j = 1; j <= 2; j++)
        for (int k = 1; k <= 2; k++) {
          callSequence[3 * level + 3] = i;
          callSequence[3 * level + 4] = j;
          callSequence[3 * level + 5] = k;
          localizeCellFindNeighbours(s, g.kid(cell, s.is[i][j][k], s.js[i][j][k], s.ks[i][j][k]), level + 1, callSequence);
        }
  } else {
    findNeighbours(s, cell, level, callSequence);
  }
}

// transportRoutinesModule.f90:1036-1054
static inline void computeCellIntensity(double& Jmean, double Iin, double Iout) {
  if (Iout < Iin) Jmean = Jmean + (Iin - Iout) / std::log(Iin / Iout);
  else Jmean = Jmean + 0.5 * (Iin + Iout);
}

// One leaf: equiSources.f90:1580-1788 (inline, unrefined base cell: `inlineBase`) and
// transportRoutinesModule.f90:588-955 (refined path).  ray ids: 0 = xy, 1 = yz, 2 = xz; the reference
// processes xy, then xz, then yz.
static int transportLeaf(Sweep& s, int cell, double weight, double cellSize, bool inlineBase) {
  Grid& g = *s.g;
  Zone& z = g.node[cell];
  const Pattern& p = s.pool[z.pattern];
  double Jm[3] = {0., 0., 0.};
  int imean = 0;
  static const int order[3] = {0, 2, 1};
  for (int q = 0; q < 3; q++) {
    int ray = order[q];
    double len;
    if (ray == 0) len = p.xy_len;
    else if (ray == 2) { if (!p.xzRayActive) continue; len = p.xz_len; }
    else { if (!p.yzRayActive) continue; len = p.yz_len; }
    double Iin[3];
    if (!z.nbPresent[ray]) {
      Iin[0] = s.uvb[0]; Iin[1] = s.uvb[1]; Iin[2] = s.uvb[2];
    } else {
      const Zone& nb = g.node[z.nb[ray]];
      const Pattern& np = s.pool[nb.pattern];
      int sel = (ray == 0) ? np.xyTop : (ray == 2 ? np.xzTop : np.yzTop);
      int src = -1;  // which of the neighbour's rays (0 xy, 1 yz, 2 xz)
      switch (sel) {
        case xyEnd: src = 0; break;
        case xzEnd:
          if (ray != 0 && !np.xzRayActive) return ERR_RAY_INACTIVE;
          src = 2; break;
        case yzEnd:
          if (ray != 0 && !np.yzRayActive) return ERR_RAY_INACTIVE;
          src = 1; break;
        default:  // selector 0
          if (inlineBase) return ERR_TOP_SELECTOR;                 // equiSources.f90:1605,1672,1741
          if (z.level <= nb.level) return ERR_TOP_SELECTOR;        // transportRoutinesModule.f90:609
          if (np.xzRayActive) {
            for (int gI = 0; gI < 3; gI++) Iin[gI] = 0.5 * (nb.Iout[2][gI] + nb.Iout[0][gI]);
          } else if (np.yzRayActive) {
            for (int gI = 0; gI < 3; gI++) Iin[gI] = 0.5 * (nb.Iout[1][gI] + nb.Iout[0][gI]);
          } else {
            for (int gI = 0; gI < 3; gI++) Iin[gI] = nb.Iout[0][gI];
          }
      }
      if (src >= 0) for (int gI = 0; gI < 3; gI++) Iin[gI] = nb.Iout[src][gI];
    }
    double dpath = cellSize * len;
    for (int gI = 0; gI < 3; gI++) {
      double tau = z.kappa[gI] * dpath;
      double tmpabs = std::exp(-tau);
      // nemi = 0: Iout = Iin*tmpabs + 0.*tmpemi/dpath (transportRoutinesModule.f90:673-678)
      double tmpemi = (tau > (double)1.e-10f) ? (1. - tmpabs) / z.kappa[gI] : dpath;
      z.Iout[ray][gI] = Iin[gI] * tmpabs + 0. * tmpemi / dpath;
    }
    if (!inlineBase) {  // guard always tests the xy ray's sum (transportRoutinesModule.f90:680,803,926)
      double tmp = z.Iout[0][0] + z.Iout[0][1] + z.Iout[0][2];
      if (!(tmp < 1.e-20 && tmp > -1.e-20)) return ERR_INTENSITY_GUARD;
    }
    for (int gI = 0; gI < 3; gI++) computeCellIntensity(Jm[gI], Iin[gI], z.Iout[ray][gI]);
    imean++;
    s.nseg++;
  }
  for (int gI = 0; gI < 3; gI++) z.Jmean[gI] = z.Jmean[gI] + Jm[gI] / (double)(float)imean * weight;
  return OK;
}

// transportRoutinesModule.f90:560-587 recursion
static int transportRec(Sweep& s, int cell, double weight, double cellSize) {
  Grid& g = *s.g;
  if (g.node[cell].refined()) {
    for (int i = 1; i <= 2; i++)
      for (int j = 1; j <= 2; j++)
        for (int k = 1; k <= 2; k++) {
          int st = transportRec(s, g.kid(cell, s.is[i][j][k], s.js[i][j][k], s.ks[i][j][k]), weight, cellSize / 2.);
          if (st) return st;
        }
    return OK;
  }
  return transportLeaf(s, cell, weight, cellSize, false);
}

// equiSources.f90:1391-1454: zone and local angles for one HEALPix direction
int directionSetup(int nAngularLevel, int64_t iray, int& izone, double& phi, double& theta) {
  int nside = 1 << (nAngularLevel - 1);
  double phiLarge, thetaLarge;
  int st = pix2ang_nest(nside, iray, phiLarge, thetaLarge);
  if (st) return st;
  double phi1, theta1;
  izone = 1;
  if (phiLarge > 0. && phiLarge < 0.5 * pi) { phi1 = phiLarge; izone += 0; }
  else if (phiLarge > 0.5 * pi && phiLarge < pi) { phi1 = phiLarge - 0.5 * pi; izone += 3; }
  else if (phiLarge > pi && phiLarge < 1.5 * pi) { phi1 = phiLarge - pi; izone += 6; }
  else if (phiLarge > 1.5 * pi && phiLarge < 2. * pi) { phi1 = phiLarge - 1.5 * pi; izone += 9; }
  else return ERR_PHI;
  if (thetaLarge > 0. && thetaLarge < 0.5 * pi) { theta1 = thetaLarge; }
  else if (thetaLarge > -0.5 * pi && thetaLarge < 0.) { theta1 = -thetaLarge; izone += 12; }
  else return ERR_THETA;
  double tmp1 = 1. / std::sin(theta1);
  double tmp2 = 1. / (std::cos(phi1) * std::cos(theta1));
  double tmp3 = 1. / (std::sin(phi1) * std::cos(theta1));
  if (tmp1 < std::fmin(tmp2, tmp3)) {
    theta = theta1; phi = phi1;
  } else if (tmp2 < std::fmin(tmp1, tmp3)) {
    theta = arcsin_(std::cos(theta1) * std::cos(phi1));
    phi = arcsin_(std::sin(theta1) / std::cos(theta));
    izone += 1;
  } else if (tmp3 < std::fmin(tmp1, tmp2)) {
    theta = arcsin_(std::cos(theta1) * std::sin(phi1));
    phi = std::acos(std::sin(theta1) / std::cos(theta));
    izone += 2;
  } else return ERR_THETA_OR_PHI;
  return OK;
}

// equiSources.f90:4956-4983
static void computeOpacities(Grid& g, const double* beta /* [3 groups][3: beta24, beta26, beta25] */) {
  for (auto& z : g.node) {
    z.Jmean[0] = z.Jmean[1] = z.Jmean[2] = 0.;
    if (!z.refined()) {
      z.kappa[0] = z.HI * beta[0];
      z.kappa[1] = z.HI * beta[3] + z.HeI * beta[4];
      z.kappa[2] = z.HI * beta[6] + z.HeI * beta[7] + z.HeII * beta[8];
    }
  }
}

// one direction: equiSources.f90:1389-1806
static int sweepDirection(Sweep& s, int nAngularLevel, int64_t iray, double weight, DiffuseTrace* tr) {
  Grid& g = *s.g;
  int nx = g.nx, ny = g.ny, nz = g.nz;
  int st = directionSetup(nAngularLevel, iray, s.izone, s.phi, s.theta);
  if (st) return st;
  int nxt, nyt, nzt;
  switch ((s.izone - 1) % 6 + 1) {  // equiSources.f90:1458-1483 (cases repeat with period 6)
    case 1: nxt = nx; nyt = ny; nzt = nz; break;
    case 2: nxt = ny; nyt = nz; nzt = nx; break;
    case 3: nxt = nz; nyt = nx; nzt = ny; break;
    case 4: nxt = nx; nyt = nz; nzt = ny; break;
    case 5: nxt = ny; nyt = nx; nzt = nz; break;
    default: nxt = nz; nyt = ny; nzt = nx; break;
  }
  for (int i = 1; i <= 2; i++)
    for (int j = 1; j <= 2; j++)
      for (int k = 1; k <= 2; k++)
        rotateIndices(i, j, k, 2, 2, 2, s.izone, s.is[i][j][k], s.js[i][j][k], s.ks[i][j][k]);
  s.pool.clear();
  s.pool.reserve((size_t)nxt * (2u << g.maxLevel) + 16);
  for (int i = 1; i <= nxt; i++) newPattern(s);
  for (int i = 1; i <= nxt; i++) {
    Pattern& p = s.pool[i - 1];
    if (i == 1) {
      p.xy_x0 = 0.5; p.xy_y0 = 0.5;
    } else {
      st = continueAbove(s.pool[i - 2], p, s.phi, s.theta);
      if (st) return st;
    }
    st = setPattern(s.pool[i - 1], s.phi, s.theta);
    if (st) return st;
    s.pool[i - 1].refined = false;
    for (int j = 1; j <= nyt; j++)
      for (int k = 1; k <= nzt; k++) {
        int ic, jc, kc;
        rotateIndices(i, j, k, nx, ny, nz, s.izone, ic, jc, kc);
        int c = g.base(ic, jc, kc);
        g.node[c].pattern = i - 1;
        if (g.node[c].refined()) {
          st = setRaysRefined(s, c, i - 1);
          if (st) return st;
        }
        g.node[c].parent = -1;
      }
  }
  int callSequence[3 * 40];
  for (int i = 1; i <= nxt; i++)
    for (int j = 1; j <= nyt; j++)
      for (int k = 1; k <= nzt; k++) {
        int ic, jc, kc;
        rotateIndices(i, j, k, nx, ny, nz, s.izone, ic, jc, kc);
        callSequence[0] = i; callSequence[1] = j; callSequence[2] = k;
        localizeCellFindNeighbours(s, g.base(ic, jc, kc), 0, callSequence);
      }
  if (tr) {
    if (tr->izoneOut) *tr->izoneOut = s.izone;
    if (tr->anglesOut) { tr->anglesOut[0] = s.phi; tr->anglesOut[1] = s.theta; }
    if (tr->patternOut)
      for (int i = 0; i < nxt; i++) {
        const Pattern& p = s.pool[i];
        double* o = tr->patternOut + 12 * i;
        o[0] = p.xy_x0; o[1] = p.xy_y0; o[2] = p.xy_len;
        o[3] = p.xzRayActive ? p.xz_x0 : 0.; o[4] = p.xzRayActive ? p.xz_z0 : 0.; o[5] = p.xzRayActive ? p.xz_len : 0.;
        o[6] = p.yzRayActive ? p.yz_y0 : 0.; o[7] = p.yzRayActive ? p.yz_z0 : 0.; o[8] = p.yzRayActive ? p.yz_len : 0.;
        o[9] = p.xyTop; o[10] = p.xzTop; o[11] = p.yzTop;
      }
    if (tr->nbLeaf) {
      int64_t nleaf = (int64_t)g.leafNode.size();
      for (int64_t l = 0; l < nleaf; l++) {
        const Zone& z = g.node[g.leafNode[l]];
        const Pattern& p = s.pool[z.pattern];
        bool act[3] = {true, p.yzRayActive, p.xzRayActive};
        for (int ray = 0; ray < 3; ray++)
          tr->nbLeaf[ray * nleaf + l] = !act[ray] ? -2 : (z.nbPresent[ray] ? g.node[z.nb[ray]].leaf : -1);
      }
    }
  }
  double cellSizeAbsoluteUnits = g.physicalBoxSize / (double)nx;
  for (int i = 1; i <= nxt; i++)
    for (int j = 1; j <= nyt; j++)
      for (int k = 1; k <= nzt; k++) {
        int ic, jc, kc;
        rotateIndices(i, j, k, nx, ny, nz, s.izone, ic, jc, kc);
        int c = g.base(ic, jc, kc);
        if (!g.node[c].refined()) st = transportLeaf(s, c, weight, cellSizeAbsoluteUnits, true);
        else st = transportRec(s, c, weight, cellSizeAbsoluteUnits);
        if (st) return st;
      }
  return OK;
}

int diffuseSolve(Grid& g, int nAngularLevel, const double* uvb, const double* beta, int64_t rayBegin, int64_t rayEnd,
                 int64_t traceRay, DiffuseTrace* tr, int64_t* nsegOut) {
  computeOpacities(g, beta);
  int64_t nrays = 12 * ((int64_t)1 << (2 * (nAngularLevel - 1)));
  double weight = (double)(1.f / (float)nrays);  // equiSources.f90:1386, single-precision division
  Sweep s;
  s.g = &g;
  s.uvb[0] = uvb[0]; s.uvb[1] = uvb[1]; s.uvb[2] = uvb[2];
  s.nseg = 0;
  if (rayBegin < 0) rayBegin = 0;
  if (rayEnd < 0 || rayEnd > nrays) rayEnd = nrays;
  for (int64_t iray = rayBegin; iray < rayEnd; iray++) {
    int st = sweepDirection(s, nAngularLevel, iray, weight, (tr && iray == traceRay) ? tr : nullptr);
    if (st) { if (nsegOut) *nsegOut = s.nseg; return st; }
  }
  if (nsegOut) *nsegOut = s.nseg;
  return OK;
}

int diffuseSolveThreaded(Grid& g, int nAngularLevel, const double* uvb, const double* beta, const int32_t* rays,
                         int nrays, int nthreads, double* J, int64_t* nsegOut) {
  computeOpacities(g, beta);
  const int64_t nraysTotal = 12 * ((int64_t)1 << (2 * (nAngularLevel - 1)));
  const double weight = (double)(1.f / (float)nraysTotal);
  const size_t nleaf = g.leafNode.size();
  if (nthreads < 1) nthreads = 1;
  const int wantCopies = nthreads - 1;
  if (nthreads > nrays) nthreads = nrays > 0 ? nrays : 1;
  std::vector<int> status(nthreads, OK);
  std::vector<int64_t> nseg(nthreads, 0);
  std::vector<std::vector<double>> part(nthreads);
  auto work = [&](int t, Grid* mine) {
    Sweep s;
    s.g = mine;
    s.uvb[0] = uvb[0]; s.uvb[1] = uvb[1]; s.uvb[2] = uvb[2];
    s.nseg = 0;
    for (int q = t; q < nrays; q += nthreads) {
      int st = sweepDirection(s, nAngularLevel, rays[q], weight, nullptr);
      if (st) { status[t] = st; break; }
    }
    nseg[t] = s.nseg;
    part[t].resize(3 * nleaf);
    for (size_t l = 0; l < nleaf; l++) {
      const Zone& z = mine->node[mine->leafNode[l]];
      for (int gI = 0; gI < 3; gI++) part[t][gI * nleaf + l] = z.Jmean[gI];
    }
  };
  // thread 0 works on g itself; the other threads on private copies that persist across calls (set-up cost, like
  // the reference's one-time octree build, is not part of the sweep)
  while ((int)g.threadCopy.size() < wantCopies) {
    Grid c;
    c.nx = g.nx; c.ny = g.ny; c.nz = g.nz; c.physicalBoxSize = g.physicalBoxSize; c.maxLevel = g.maxLevel;
    c.node = g.node; c.leafNode = g.leafNode;
    g.threadCopy.push_back(std::move(c));
  }
  for (int t = 1; t < nthreads; t++) {  // refresh what may have changed since the copy: species -> kappa, J = 0
    Grid& c = g.threadCopy[t - 1];
    for (size_t i = 0; i < g.node.size(); i++) {
      c.node[i].HI = g.node[i].HI; c.node[i].HeI = g.node[i].HeI; c.node[i].HeII = g.node[i].HeII;
    }
    computeOpacities(c, beta);
  }
  std::vector<std::thread> th;
  for (int t = 1; t < nthreads; t++) th.emplace_back(work, t, &g.threadCopy[t - 1]);
  work(0, &g);
  for (auto& x : th) x.join();
  int64_t tot = 0;
  for (size_t i = 0; i < 3 * nleaf; i++) {
    double sum = 0.;
    for (int t = 0; t < nthreads; t++) sum += part[t][i];
    J[i] = sum;
  }
  for (int t = 0; t < nthreads; t++) { tot += nseg[t]; if (status[t]) return status[t]; }
  if (nsegOut) *nsegOut = tot;
  return OK;
}

}  // namespace ftte
