// Host-side direction / pattern tables (see geometry.h).  Compiled without FMA contraction so that every
// comparison below sees the same doubles the reference's gfortran -O2 build would produce.
#include "geometry.h"

#include <cmath>

#include "../../include/rtb200.h"

namespace rtb {

// ---------------------------------------------------------------------------------------------------
// 24-zone index rotation as a table.  Zones come in blocks of three (dominant component = rotated i maps to
// physical x, y, z); blocks 2-4 add reflections of the two transverse axes, zones 13-24 reflect the first
// physical component on top (rotateIndicesModule.f90:14-111).
// ---------------------------------------------------------------------------------------------------
ZoneMap zone_map(int izone) {
  static const int8_t kSrc[12][3] = {
      {0, 1, 2}, {1, 2, 0}, {2, 0, 1},   // 1-3
      {0, 2, 1}, {1, 0, 2}, {2, 1, 0},   // 4-6  (third component reflected)
      {0, 1, 2}, {1, 2, 0}, {2, 0, 1},   // 7-9  (second and third reflected)
      {0, 2, 1}, {1, 0, 2}, {2, 1, 0}};  // 10-12 (second reflected)
  static const int8_t kRefl[4][3] = {{0, 0, 0}, {0, 0, 1}, {0, 1, 1}, {0, 1, 0}};
  ZoneMap m;
  int b = (izone - 1) % 12;
  for (int c = 0; c < 3; c++) {
    m.src[c] = kSrc[b][c];
    m.refl[c] = kRefl[b / 3][c];
  }
  if (izone > 12) m.refl[0] = 1;
  return m;
}

ZoneStrides zone_strides(int izone, int n) {
  const int64_t phys[3] = {(int64_t)n * n, n, 1};
  return zone_strides_layout(izone, n, phys);
}

ZoneStrides zone_strides_layout(int izone, int n, const int64_t phys[3]) {
  ZoneMap m = zone_map(izone);
  ZoneStrides s;
  s.origin = 0;
  for (int c = 0; c < 3; c++) {
    int r = m.src[c];
    if (m.refl[c]) {
      s.stride[r] = -phys[c];
      s.origin += (int64_t)(n - 1) * phys[c];
    } else {
      s.stride[r] = phys[c];
    }
  }
  return s;
}

// ---------------------------------------------------------------------------------------------------
// HEALPix NESTED pixel centre, in the reference's conventions: theta is the latitude acos(z) - pi/2 and the
// result is rotated by 0.111 rad about x and 0.222 rad about y (equiSources.f90:2118-2231, 2297-2361).
// ---------------------------------------------------------------------------------------------------
static inline double clamped_asin(double x) {
  return x > 1.0 ? kHalfPi : (x < -1.0 ? -kHalfPi : std::asin(x));
}

static inline double angle_from(double c, double s) {
  double a = clamped_asin(s);
  if (c > 0.) return s > 0. ? a : kTwoPi + a;
  return kPi - a;
}

static inline int deinterleave_even(int v) {  // bits 0,2,4,... of v packed together (mk_pix2xy)
  int r = 0;
  for (int b = 0; b < 5; b++) r |= ((v >> (2 * b)) & 1) << b;
  return r;
}

static void tilt(double& phi, double& theta) {
  const double a1 = (double)0.111f, a2 = (double)0.222f;
  double p0 = phi, t0 = theta;
  theta = clamped_asin(std::cos(t0) * std::sin(p0) * std::sin(a1) + std::sin(t0) * std::cos(a1));
  double c = std::cos(t0) * std::cos(p0) / std::cos(theta);
  double s = (std::cos(t0) * std::sin(p0) * std::cos(a1) - std::sin(t0) * std::sin(a1)) / std::cos(theta);
  phi = angle_from(c, s);
  p0 = phi; t0 = theta;
  theta = clamped_asin(std::cos(t0) * std::cos(p0) * std::sin(a2) + std::sin(t0) * std::cos(a2));
  c = (std::cos(t0) * std::cos(p0) * std::cos(a2) - std::sin(t0) * std::sin(a2)) / std::cos(theta);
  s = std::cos(t0) * std::sin(p0) / std::cos(theta);
  phi = angle_from(c, s);
}

int healpix_center(int nside, int64_t ipix, double* phiOut, double* thetaOut) {
  static const int kRing[12] = {2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4};
  static const int kPhase[12] = {1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7};
  if (nside < 1 || nside > 32768) return RTB200_ERR_ARG;
  const int64_t perFace = (int64_t)nside * nside;
  if (ipix < 0 || ipix >= 12 * perFace) return RTB200_ERR_ARG;
  const int face = (int)(ipix / perFace);
  const int64_t inFace = ipix % perFace;
  int ix = 0, iy = 0;
  for (int chunk = 0; chunk < 3; chunk++) {  // 10 bits at a time, weights 1, 32, 1024
    int bits = (int)((inFace >> (10 * chunk)) & 1023);
    ix += deinterleave_even(bits) << (5 * chunk);
    iy += deinterleave_even(bits >> 1) << (5 * chunk);
  }
  const double fn = (double)(float)nside;
  const double fact1 = 1.0 / (3.0 * fn * fn), fact2 = 2.0 / (3.0 * fn);
  const int nl4 = 4 * nside;
  const int ring = kRing[face] * nside - (ix + iy) - 1;
  int nr = nside, shift = (ring - nside) % 2;
  double z = (double)(float)(2 * nside - ring) * fact2;
  if (ring < nside) {
    nr = ring; shift = 0;
    z = 1.0 - (double)((float)nr * (float)nr) * fact1;
  } else if (ring > 3 * nside) {
    nr = nl4 - ring; shift = 0;
    z = -1.0 + (double)((float)nr * (float)nr) * fact1;
  }
  double theta = std::acos(z) - kHalfPi;
  int jp = (kPhase[face] * nr + (ix - iy) + 1 + shift) / 2;
  if (jp > nl4) jp -= nl4;
  if (jp < 1) jp += nl4;
  double phi = (double)((float)jp - (float)(shift + 1) * 0.5f) * kHalfPi / (double)(float)nr;
  while (phi > kTwoPi) phi -= kTwoPi;
  while (phi < 0.) phi += kTwoPi;
  tilt(phi, theta);
  if (phi > 2.0 * kPi) return RTB200_ERR_ANGLE_LARGE;
  *phiOut = phi;
  *thetaOut = theta;
  return RTB200_OK;
}

// zone = 1 + 3*quadrant(phi) + 12*[theta<0] + dominant component; local angles measured from the dominant
// axis so that 1/sin(theta) is the shortest exit (equiSources.f90:1395-1454).
Direction classify_direction(int nAngularLevel, int64_t iray) {
  Direction d;
  d.iray = iray; d.izone = 0; d.phi = d.theta = 0.;
  double P, T;
  d.status = healpix_center(1 << (nAngularLevel - 1), iray, &P, &T);
  if (d.status) return d;
  int zone = 1;
  double p1, t1;
  if (P > 0. && P < 0.5 * kPi) { p1 = P; }
  else if (P > 0.5 * kPi && P < kPi) { p1 = P - 0.5 * kPi; zone += 3; }
  else if (P > kPi && P < 1.5 * kPi) { p1 = P - kPi; zone += 6; }
  else if (P > 1.5 * kPi && P < 2. * kPi) { p1 = P - 1.5 * kPi; zone += 9; }
  else { d.status = RTB200_ERR_PHI; return d; }
  if (T > 0. && T < 0.5 * kPi) { t1 = T; }
  else if (T > -0.5 * kPi && T < 0.) { t1 = -T; zone += 12; }
  else { d.status = RTB200_ERR_THETA; return d; }
  const double e1 = 1. / std::sin(t1);
  const double e2 = 1. / (std::cos(p1) * std::cos(t1));
  const double e3 = 1. / (std::sin(p1) * std::cos(t1));
  if (e1 < std::fmin(e2, e3)) {
    d.theta = t1; d.phi = p1;
  } else if (e2 < std::fmin(e1, e3)) {
    d.theta = clamped_asin(std::cos(t1) * std::cos(p1));
    d.phi = clamped_asin(std::sin(t1) / std::cos(d.theta));
    zone += 1;
  } else if (e3 < std::fmin(e1, e2)) {
    d.theta = clamped_asin(std::cos(t1) * std::sin(p1));
    d.phi = std::acos(std::sin(t1) / std::cos(d.theta));
    zone += 2;
  } else { d.status = RTB200_ERR_THETA_OR_PHI; return d; }
  d.izone = zone;
  return d;
}

// ---------------------------------------------------------------------------------------------------
// Segment geometry of one layer given where its xy ray enters the bottom face
// (transportRoutinesModule.f90:7-85; case table in SURVEY.md appendix A).
// ---------------------------------------------------------------------------------------------------
static void solve_layer(RayPattern& p, double phi, double theta) {
  const double st = std::sin(theta), ct = std::cos(theta), sp = std::sin(phi), cp = std::cos(phi);
  const double tTop = 1. / st;
  const double tX = (1. - p.xy_x0) / (cp * ct);
  const double tY = (1. - p.xy_y0) / (sp * ct);
  p.xzActive = p.yzActive = 0;
  p.xzTop = p.yzTop = 0;
  p.status = RTB200_OK;
  if (tTop < std::fmin(tX, tY)) {           // straight through the top
    p.xy_len = tTop;
    p.xyTop = 1;
  } else if (tX < std::fmin(tTop, tY)) {    // leaves through x = 1, continues next door as a yz ray
    p.xy_len = tX;
    p.yzActive = 1;
    p.yz_y0 = (1. - p.xy_x0) * std::tan(phi) + p.xy_y0;
    p.yz_z0 = p.xy_len * st;
    if (p.yz_y0 > 1. || p.yz_z0 > 1.) { p.status = RTB200_ERR_PATTERN_RANGE; return; }
    const double a1 = (1. - p.yz_z0) / st;
    const double a2 = (1. - p.yz_y0) / (sp * ct);
    p.yzTop = 1;
    if (a1 < a2) {
      p.yz_len = a1;
      p.xyTop = 2;
    } else {
      p.yz_len = a2;
      p.xzActive = 1;
      p.xz_x0 = (1. - p.yz_y0) / std::tan(phi);
      p.xz_z0 = p.yz_z0 + a2 * st;
      p.xz_len = (1. - p.xz_z0) / st;
      p.xyTop = 3; p.xzTop = 2;
    }
  } else {                                   // leaves through y = 1, continues as an xz ray
    p.xy_len = tY;
    p.xzActive = 1;
    p.xz_x0 = (1. - p.xy_y0) / std::tan(phi) + p.xy_x0;
    p.xz_z0 = tY * st;
    if (p.xz_x0 > 1. || p.xz_z0 > 1.) { p.status = RTB200_ERR_PATTERN_RANGE; return; }
    const double b1 = (1. - p.xz_z0) / st;
    const double b2 = (1. - p.xz_x0) / (cp * ct);
    p.xzTop = 1;
    if (b1 < b2) {
      p.xz_len = b1;
      p.xyTop = 3;
    } else {
      p.xz_len = b2;
      p.yzActive = 1;
      p.yz_y0 = (1. - p.xz_x0) * std::tan(phi);
      p.yz_z0 = p.xz_len * st + p.xz_z0;
      p.yz_len = (1. - p.yz_z0) / st;
      p.xyTop = 2; p.yzTop = 3;
    }
  }
}

// where the ray that leaves `below` through its top face enters the next layer
// (equiSources.f90:1507-1522, transportRoutinesModule.f90:167-182)
static void enter_from_below(const RayPattern& below, RayPattern& cur, double phi, double theta) {
  switch (below.xyTop) {
    case 1:
      cur.xy_x0 = below.xy_x0 + std::cos(phi) / std::tan(theta);
      cur.xy_y0 = below.xy_y0 + std::sin(phi) / std::tan(theta);
      break;
    case 3:
      cur.xy_x0 = below.xz_x0 + below.xz_len * std::cos(theta) * std::cos(phi);
      cur.xy_y0 = below.xz_len * std::cos(theta) * std::sin(phi);
      break;
    case 2:
      cur.xy_x0 = below.yz_len * std::cos(theta) * std::cos(phi);
      cur.xy_y0 = below.yz_y0 + below.yz_len * std::cos(theta) * std::sin(phi);
      break;
    default:
      cur.status = RTB200_ERR_TOP_SELECTOR;
      return;
  }
  cur.status = (cur.xy_x0 > 1. || cur.xy_y0 > 1.) ? RTB200_ERR_PATTERN_RANGE : RTB200_OK;
}

static RayPattern blank() {
  RayPattern p;
  p.xy_x0 = p.xy_y0 = p.xy_len = p.xz_x0 = p.xz_z0 = p.xz_len = p.yz_y0 = p.yz_z0 = p.yz_len = 0.;
  p.xzActive = p.yzActive = p.xyTop = p.xzTop = p.yzTop = 0;
  p.status = RTB200_OK;
  return p;
}

void layer_patterns_level0(double phi, double theta, int n, std::vector<RayPattern>& out) {
  out.assign(n, blank());
  for (int i = 0; i < n; i++) {
    RayPattern& p = out[i];
    if (i == 0) {
      p.xy_x0 = 0.5; p.xy_y0 = 0.5;
    } else {
      if (out[i - 1].status) { p.status = out[i - 1].status; continue; }
      enter_from_below(out[i - 1], p, phi, theta);
      if (p.status) continue;
    }
    solve_layer(p, phi, theta);
  }
}

// A refined cell splits its layer in two: the lower sub-layer starts at frac(2*x0), frac(2*y0) of the parent
// entry, the upper one where the lower one's top-leaving ray arrives (transportRoutinesModule.f90:151-187).
void layer_patterns_refine(double phi, double theta, const std::vector<RayPattern>& parent,
                           std::vector<RayPattern>& out) {
  out.assign(parent.size() * 2, blank());
  for (size_t i = 0; i < parent.size(); i++) {
    const RayPattern& pp = parent[i];
    RayPattern& lo = out[2 * i];
    RayPattern& up = out[2 * i + 1];
    if (pp.status) { lo.status = up.status = pp.status; continue; }
    lo.xy_x0 = pp.xy_x0 < 0.5 ? 2. * pp.xy_x0 : 2. * pp.xy_x0 - 1.;
    lo.xy_y0 = pp.xy_y0 < 0.5 ? 2. * pp.xy_y0 : 2. * pp.xy_y0 - 1.;
    solve_layer(lo, phi, theta);
    if (lo.status) { up.status = lo.status; continue; }
    enter_from_below(lo, up, phi, theta);
    if (up.status) continue;
    solve_layer(up, phi, theta);
  }
}

}  // namespace rtb
