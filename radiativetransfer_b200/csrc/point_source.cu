// Point-source ray casting with per-cell rate deposition on the GPU.
//
// Replaces the reference's source loop and its internal procedures (all in equiSources.f90):
//   :1256-1370  loop over sources, 12 base rays each          -> point_solve() below
//   :3120-3385  startNewLongRay (march / deposit / split x4)   -> point_march_kernel, one launch per pixel level
//   :2412-2595  drawSegment                                    -> draw_segment()
//   :2647-2960  find/zoom{XY,YZ,XZ}Neighbour                   -> step_neighbour() on the linear octree
//   :3011-3118  absoluteCoordinates, localizeSplitContinuationCell -> child_start()
//   :4157-4311  getRatesHydrogenHelium                         -> rates_faithful() / FAST-mode slopes
//   stellarBetaTable.f90:217-285 (400 x 11^4 table sums)       -> point_table_kernel (stores LOG tables)
//
// Formulation.  The reference recurses depth-first: a ray marches until its radius reaches rmax(pixelLevel), then
// splits into the 4 nested HEALPix children, which inherit the accumulated optical depths.  Here the ray tree is
// processed breadth-first, one kernel launch per pixel level: thread = one ray object (source, pixel); it reads its
// parent's end state (leaf, point, radius, depths), applies the reference's continuation arithmetic, marches and
// deposits, and stores its own end state for the next level.  The arithmetic of every ray is that of the
// reference, operation for operation (no FMA contraction in anything that decides a branch); only the order in which
// different rays add to the same cell differs.
//
// Data layout: the grid is the per-leaf SoA of the context (leaf order of the reference) + the linear octree
// `child[]`; per-source log-tables [planes][11^3][8 slots] (85 KB without dust: L1/L2 resident); ray end states
// [source][pixel] AoS of 80 B, ping-pong between levels; rates [6][nleaf] fp64, accumulated with fp64 RED
// (atomicAdd, native on sm_100) -- the deterministic, atomic-free segmented variant sorts (leaf, deposit) records.
#include <algorithm>
#include <cmath>
#include <cstdio>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#include "point_host.h"
#include "portable_math.h"
#include "rtb200_internal.h"
#include "segment_math.cuh"

namespace rtb {

namespace {

constexpr int kPlane = 11 * 11 * 11;
// log-tables are node-major: [source][dust plane][i3][i2][i1][kSlots], slot 2r = number table of reaction r (24, 26,
// 25), slot 2r+1 = its energy table, slots 6,7 unused: a node is one 64-byte record, and the (number, energy) pair of
// a reaction is one 16-byte load
constexpr int kSlots = 8;
constexpr int kDiagStride = 320;  // per source: remaining[7], boundary[7], dust, highestPixelLevel, spectrum[300]
constexpr int kMaxPixelLevel = 8;

__device__ __forceinline__ double M(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double A(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double S(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double D(double a, double b) { return __ddiv_rn(a, b); }

struct RayState {  // end state of a ray object, read by its 4 children
  double x, y, z, radius, d1, d2, d3, dD;
  int32_t leaf;
  int32_t strategy;  // 2 = split (children continue), anything else: children do not exist
  int32_t pad[2];
};
static_assert(sizeof(RayState) == 80, "RayState layout");

struct PointParams {
  // grid
  const int32_t* child;
  const int8_t* level;
  const int32_t *leafX, *leafY, *leafZ;
  const double *HI, *HeI, *HeII, *rho, *abun2;
  int64_t nleaf;
  int nx;
  double boxSize;
  // sources of this batch
  const int32_t* srcLeaf;
  const int32_t* srcWeight;
  const double* logTab;   // [nsrc][planes][kPlane][kSlots]
  int planes;             // 1 without dust, 11 with
  int dust;
  int maxPixelLevel;
  // tables
  const double* pixDir;   // [npix(levels 1..max)][3]
  double rmax[kMaxPixelLevel + 2];
  double outRadiusKpc[7]; // outputRadius(i) * kpc  [cm]
  double outRadius[7];
  double kpc;
  double outRadiusCells[7];  // outputRadius(i)*kpc in base-cell units: a conservative pre-filter for the exact test
  double cellSize[32];       // physicalBoxSize / (float(2**level) * float(nx)), equiSources.f90:3176
  const double* outSigma; // [4][300] ratios sigma(nu)/sigma_threshold: 24, 26, 25, dust
  // outputs
  double* rates;          // [6][nleaf]: krate24, krate25, krate26, crate24, crate25, crate26
  double* diag;           // [nsrc][kDiagStride]
  RayState* stateIn;      // level-1 states  [nsrc][npix(level-1)]
  RayState* stateOut;     // this level's    [nsrc][npix(level)]
  unsigned long long* nseg;
  int32_t* err;
  // segmented (atomic-free) deposition: records instead of RED.ADD, see segmented_reduce_kernel
  long long* recKey;      // [recCap] leaf << 32 | ray << 12 | segment
  double* recVal;         // [6][recCap]
  unsigned long long* recCount;
  long long recCap;
  int raysPerSource;      // pixels of levels 1..maxPixelLevel
  // planned deposition (deposit mode 2): the ray geometry of a source batch is cached, every segment of every ray has a
  // fixed slot in a leaf-ordered record array
  unsigned int* rayCount;       // plan pass only: [rays of the batch] segments of every ray object
  const long long* rayBase;     // [rays of the batch] first index of the ray's segments in slotMap
  const unsigned int* slotMap;  // [records] (rayBase[ray] + segment) -> position in leaf order
  double* recVal6;              // [records][8] deposits in leaf order: six values + pad = one aligned 64-byte DRAM atom
  int* queue;             // [nsrc] next pixel of the last level to hand out (NULL: static pixel -> thread map)
  // optional traversal trace (parity checks)
  long long* trace;       // [cap][2]
  unsigned long long* traceLen;
  long long traceCap;
};

struct Cell {
  int32_t leaf, lvl, X, Y, Z;  // integer coordinates at the leaf's own level
};

__device__ __forceinline__ int64_t pix_offset(int level) {  // first pixel of `level` in the direction table
  return 4 * ((1LL << (2 * (level - 1))) - 1);              // 12 * (4^(L-1) - 1) / 3
}

// find??Neighbour + zoom??Neighbour (equiSources.f90:2647-2960).  axis = normal of the exit face (0 x, 1 y, 2 z);
// (a, b) = the exit point in the two in-face coordinates, ordered (x,y) / (y,z) / (x,z).  Climb while the cell lies on
// that face of its parent (halving the in-face coordinates), step to the sibling container, descend with the
// reference's `.lt.0.5` tests.  Returns false at the domain boundary.
__device__ __forceinline__ bool step_neighbour(const PointParams& P, Cell& c, int axis, int side, double& a, double& b) {
  const int L = c.lvl;
  const int Cax = axis == 0 ? c.X : (axis == 1 ? c.Y : c.Z);
  const int CA = axis == 2 ? c.X : (axis == 0 ? c.Y : c.X);
  const int CB = axis == 2 ? c.Y : c.Z;
  int l = L;
  while (l > 0) {
    const int sh = L - l;
    if (((Cax >> sh) & 1) != side) break;
    a = ((CA >> sh) & 1) ? A(M(0.5, a), 0.5) : M(0.5, a);
    b = ((CB >> sh) & 1) ? A(M(0.5, b), 0.5) : M(0.5, b);
    l--;
  }
  const int sh = L - l;
  int cx = c.X >> sh, cy = c.Y >> sh, cz = c.Z >> sh;
  if (l == 0) {
    const int base = axis == 0 ? cx : (axis == 1 ? cy : cz);
    if ((side == 0 && base == 0) || (side == 1 && base == P.nx - 1)) return false;
  }
  const int stepv = side ? 1 : -1;
  if (axis == 0) cx += stepv; else if (axis == 1) cy += stepv; else cz += stepv;
  int node = ((cx >> l) * P.nx + (cy >> l)) * P.nx + (cz >> l);
  for (int t = l - 1; t >= 0; t--)
    node = __ldg(P.child + node) + ((((cx >> t) & 1) << 2) | (((cy >> t) & 1) << 1) | ((cz >> t) & 1));
  int ch;
  while ((ch = __ldg(P.child + node)) >= 0) {
    int ia, ib;
    if (a < 0.5) { a = M(2., a); ia = 0; } else { a = S(M(2., a), 1.); ia = 1; }
    if (b < 0.5) { b = M(2., b); ib = 0; } else { b = S(M(2., b), 1.); ib = 1; }
    const int in = side == 0 ? 1 : 0;
    int bx, by, bz;
    if (axis == 2) { bx = ia; by = ib; bz = in; }
    else if (axis == 0) { bx = in; by = ia; bz = ib; }
    else { bx = ia; by = in; bz = ib; }
    node = ch + ((bx << 2) | (by << 1) | bz);
    cx = 2 * cx + bx; cy = 2 * cy + by; cz = 2 * cz + bz;
    l++;
  }
  c.leaf = -ch - 1; c.lvl = l; c.X = cx; c.Y = cy; c.Z = cz;
  return true;
}

// --- table lookup, reference operation sequence (equiSources.f90:4205-4238) -----------------------------------------
template <bool PORTABLE>
__device__ __forceinline__ double exp_ref(double x) { return PORTABLE ? rtb_pm::pm_exp(x) : exp(x); }

struct Pair {
  double n, e;  // number table, energy table
};
__device__ __forceinline__ Pair ld_pair(const double2* p) {
  const double2 v = __ldg(p);
  return Pair{v.x, v.y};
}

// both tables of one reaction at once: T points at slot 2r of node (0,0,0) of the dust plane; node stride = kSlots
// doubles = 4 double2.  Each table's operation sequence is the reference's (equiSources.f90:4205-4238).
__device__ __forceinline__ Pair interp_log(const double2* __restrict__ T, int i1, int i2, int i3, double c1, double c2,
                                           double c3) {
  constexpr int N = kSlots / 2;
  const double m3 = S(1., c3), m2 = S(1., c2), m1 = S(1., c1);
  const double w00 = M(m3, m2), w01 = M(c3, m2), w10 = M(c2, m3), w11 = M(c3, c2);
  const double2* p = T + ((i3 * 11 + i2) * 11 + i1) * N;
  const Pair u0 = ld_pair(p + N), u1 = ld_pair(p + 122 * N), u2 = ld_pair(p + 12 * N), u3 = ld_pair(p + 133 * N);
  const Pair l0 = ld_pair(p), l1 = ld_pair(p + 121 * N), l2 = ld_pair(p + 11 * N), l3 = ld_pair(p + 132 * N);
  Pair r;
  {
    const double up = A(A(A(M(w00, u0.n), M(w01, u1.n)), M(w10, u2.n)), M(w11, u3.n));
    const double lo = A(A(A(M(w00, l0.n), M(w01, l1.n)), M(w10, l2.n)), M(w11, l3.n));
    r.n = A(M(c1, up), M(m1, lo));
  }
  {
    const double up = A(A(A(M(w00, u0.e), M(w01, u1.e)), M(w10, u2.e)), M(w11, u3.e));
    const double lo = A(A(A(M(w00, l0.e), M(w01, l1.e)), M(w10, l2.e)), M(w11, l3.e));
    r.e = A(M(c1, up), M(m1, lo));
  }
  return r;
}

struct DepthIdx {
  int i1, i2, i3, iD;
  double c1, c2, c3, cD;
  int status;  // 0 ok, 1 beyond the table (rates are zero), <0 error
};

__device__ __forceinline__ DepthIdx depth_index(double t1, double t2, double t3, double tD, int dust) {
  DepthIdx q;
  q.status = 0;
  if (t1 > 10. || t2 > 10. || t3 > 10. || tD > 10.) { q.status = 1; return q; }
  q.i1 = (int)M(D(t1, 10.), 10.); q.i2 = (int)M(D(t2, 10.), 10.); q.i3 = (int)M(D(t3, 10.), 10.);
  q.c1 = S(D(M(t1, 10.), 10.), (double)q.i1);
  q.c2 = S(D(M(t2, 10.), 10.), (double)q.i2);
  q.c3 = S(D(M(t3, 10.), 10.), (double)q.i3);
  if (dust == 0) { q.iD = 0; q.cD = 0.; }
  else { q.iD = (int)M(D(tD, 10.), 10.); q.cD = S(D(M(tD, 10.), 10.), (double)q.iD); }
  if (min(min(q.i1, q.i2), min(q.i3, q.iD)) < 0) q.status = -1;
  // tau == 10 exactly: the reference indexes one past its (0:10) tables; reported instead of reproduced
  if (q.i1 >= 10 || q.i2 >= 10 || q.i3 >= 10 || q.iD >= 10) q.status = -1;
  return q;
}

// FAST mode: index = floor(tau) (table spacing 10/10 = 1) without the reference's tau/10.*10. round trip; the two can
// differ only where tau is within an ulp of a node, where the interpolant is continuous
__device__ __forceinline__ DepthIdx depth_index_fast(double t1, double t2, double t3, double tD, int dust) {
  DepthIdx q;
  q.status = 0;
  if (t1 > 10. || t2 > 10. || t3 > 10. || tD > 10.) { q.status = 1; return q; }
  q.i1 = min((int)t1, 9); q.i2 = min((int)t2, 9); q.i3 = min((int)t3, 9);
  q.c1 = t1 - (double)q.i1; q.c2 = t2 - (double)q.i2; q.c3 = t3 - (double)q.i3;
  if (dust == 0) { q.iD = 0; q.cD = 0.; }
  else { q.iD = min((int)tD, 9); q.cD = tD - (double)q.iD; }
  if (min(min(q.i1, q.i2), min(q.i3, q.iD)) < 0) q.status = -1;
  return q;
}

template <bool PORTABLE>
__device__ __forceinline__ int rates_faithful(const PointParams& P, const double* __restrict__ LT, int r, double t1,
                                              double t2, double t3, double tD, double& num, double& heat) {
  const DepthIdx q = depth_index(t1, t2, t3, tD, P.dust);
  if (q.status == 1) { num = 0.; heat = 0.; return 0; }
  if (q.status < 0) return RTB200_ERR_IDEPTH;
  const double2* T = reinterpret_cast<const double2*>(LT + (size_t)q.iD * kPlane * kSlots + 2 * r);
  const Pair v1 = interp_log(T, q.i1, q.i2, q.i3, q.c1, q.c2, q.c3);
  if (P.dust) {
    const Pair v2 = interp_log(T + kPlane * (kSlots / 2), q.i1, q.i2, q.i3, q.c1, q.c2, q.c3);
    const double mD = S(1., q.cD);
    num = exp_ref<PORTABLE>(A(M(mD, v1.n), M(q.cD, v2.n)));
    heat = exp_ref<PORTABLE>(A(M(mD, v1.e), M(q.cD, v2.e)));
  } else {  // cDust = 0: exp((1.-0.)*nr1 + 0.*nr2) == exp(nr1) exactly; the second dust plane is never needed
    num = exp_ref<PORTABLE>(v1.n);
    heat = exp_ref<PORTABLE>(v1.e);
  }
  return 0;
}

// --- FAST mode: the same piecewise-multilinear log-interpolant, evaluated without the cancellation -----------------
// R_r(d) - R_r(d + tau e_r) = R_r(d) * (1 - exp(dlog)),  dlog = nr(d + tau e_r) - nr(d) = integral of the interpolant's
// slope along axis r, which is piecewise constant: per crossed table cell the slope is a bilinear (trilinear with dust)
// combination of node differences.  One exponential + one expm1 per (reaction, table) instead of two exponentials whose
// difference cancels, and 8 table reads instead of 16 (32 with dust): the start value and the slope share their nodes.
// value of the log-interpolant restricted to the table nodes with index iAx along AXIS (0,1,2 = tau1,tau2,tau3), at
// the transverse position of q: a bilinear (with dust: trilinear) combination of 4 (8) nodes
template <int AXIS>
__device__ __forceinline__ Pair face_sum(const double2* __restrict__ T, int iAx, const DepthIdx& q, int dust) {
  constexpr int N = kSlots / 2;                                        // node stride in double2
  constexpr int sA = (AXIS == 0 ? 1 : (AXIS == 1 ? 11 : 121)) * N;
  constexpr int s1 = (AXIS == 0 ? 11 : 1) * N, s2 = (AXIS == 2 ? 11 : 121) * N;   // the two transverse depth axes
  const int j1 = AXIS == 0 ? q.i2 : q.i1, j2 = AXIS == 2 ? q.i2 : q.i3;
  const double c1 = AXIS == 0 ? q.c2 : q.c1, c2 = AXIS == 2 ? q.c2 : q.c3;
  const double2* p = T + iAx * sA + j1 * s1 + j2 * s2 + q.iD * (kPlane * N);
  const Pair a0 = ld_pair(p), a1 = ld_pair(p + s1), b0 = ld_pair(p + s2), b1 = ld_pair(p + s1 + s2);
  Pair v;
  {
    const double lo = fma(c1, a1.n - a0.n, a0.n), hi = fma(c1, b1.n - b0.n, b0.n);
    v.n = fma(c2, hi - lo, lo);
  }
  {
    const double lo = fma(c1, a1.e - a0.e, a0.e), hi = fma(c1, b1.e - b0.e, b0.e);
    v.e = fma(c2, hi - lo, lo);
  }
  if (dust) {
    const double2* d = p + kPlane * N;
    const Pair e0 = ld_pair(d), e1 = ld_pair(d + s1), f0 = ld_pair(d + s2), f1 = ld_pair(d + s1 + s2);
    {
      const double lo2 = fma(c1, e1.n - e0.n, e0.n), hi2 = fma(c1, f1.n - f0.n, f0.n);
      const double v2 = fma(c2, hi2 - lo2, lo2);
      v.n = fma(q.cD, v2 - v.n, v.n);
    }
    {
      const double lo2 = fma(c1, e1.e - e0.e, e0.e), hi2 = fma(c1, f1.e - f0.e, f0.e);
      const double v2 = fma(c2, hi2 - lo2, lo2);
      v.e = fma(q.cD, v2 - v.e, v.e);
    }
  }
  return v;
}

// deposit of one reaction (number and energy) for a step tau along its own depth axis, FAST mode.  Along that axis the
// interpolant is piecewise linear with nodes at the integers (table spacing 10/10 = 1), so
//   log R(end) - log R(start) = sum over the crossed table cells of  (length inside the cell) * (H - L),
// L/H = face_sum at the cell's lower/upper node; the upper face of one cell is the lower face of the next.
// e^x for |x| <= ~700 and 1 - e^-t for t >= 0 from the sweep's table-based exponential (segment_math.cuh: 16-entry
// shared-memory table x degree-7 polynomial, ~1e-16 relative): 14 and 15 instructions instead of libm's ~45
__device__ __forceinline__ double exp_tab(double x, const double* __restrict__ sT) {
  double Ts, p;
  exp_neg_parts<false>(-x, sT, Ts, p);
  return fma(Ts, p, Ts);
}
__device__ __forceinline__ double one_minus_exp_neg_tab(double t, const double* __restrict__ sT) {
  double Ts, p;
  exp_neg_parts<true>(t, sT, Ts, p);
  return fma(-Ts, p, 1.0 - Ts);   // exact 1 - Ts, and for t < ln2/32 simply -p: no cancellation
}

template <int AXIS>
__device__ __forceinline__ void rates_fast(const PointParams& P, const double* __restrict__ LT, const DepthIdx& q,
                                           double depthAxis, double tau, const double* __restrict__ sT, double& dnum,
                                           double& dheat) {
  const double2* T = reinterpret_cast<const double2*>(LT + 2 * AXIS);   // slot pair of this reaction
  if (tau == 0.) { dnum = 0.; dheat = 0.; return; }  // R(d) - R(d) (e.g. no helium: tau2 = tau3 = 0)
  const int i0 = AXIS == 0 ? q.i1 : (AXIS == 1 ? q.i2 : q.i3);
  const double c0 = AXIS == 0 ? q.c1 : (AXIS == 1 ? q.c2 : q.c3);
  Pair L = face_sum<AXIS>(T, i0, q, P.dust), H = face_sum<AXIS>(T, i0 + 1, q, P.dust);
  const double n0 = exp_tab(fma(c0, H.n - L.n, L.n), sT), h0 = exp_tab(fma(c0, H.e - L.e, L.e), sT);
  const double end = depthAxis + tau;
  if (end > 10.) { dnum = n0; dheat = h0; return; }  // the table returns 0 beyond tau = 10
  int iEnd = (int)end;
  if (iEnd > 9) iEnd = 9;
  double dn, dh;
  if (iEnd <= i0) {  // the step stays inside the start cell (the common case)
    dn = tau * (H.n - L.n);
    dh = tau * (H.e - L.e);
  } else {
    const double first = (double)(i0 + 1) - depthAxis;
    dn = first * (H.n - L.n);
    dh = first * (H.e - L.e);
    for (int c = i0 + 1; c <= iEnd; c++) {
      L = H;
      H = face_sum<AXIS>(T, c + 1, q, P.dust);
      const double len = c == iEnd ? end - (double)c : 1.0;
      dn = fma(len, H.n - L.n, dn);
      dh = fma(len, H.e - L.e, dh);
    }
  }
  // the tables fall with depth (dn, dh <= 0); a rising table (never with physical spectra) takes the libm route
  dnum = dn <= 0. ? n0 * one_minus_exp_neg_tab(-dn, sT) : -n0 * expm1(dn);
  dheat = dh <= 0. ? h0 * one_minus_exp_neg_tab(-dh, sT) : -h0 * expm1(dh);
}

// --------------------------------------------------------------------------------------------------------------------
// DEPOSIT: 0 = fp64 RED.ADD into the rate fields, 1 = (key, 6 deposits) records for the sort, 2 = planned slots
template <bool FAITHFUL, bool PORTABLE, bool TRACE, int DEPOSIT, int MINB = 4>
__global__ void __launch_bounds__(128, MINB) point_march_kernel(const __grid_constant__ PointParams P, int pixelLevel) {
  __shared__ double sT[16];
  if (threadIdx.x < 16) sT[threadIdx.x] = kExpTable[threadIdx.x];
  __syncthreads();
  const int64_t npix = 12LL << (2 * (pixelLevel - 1));
  const int64_t ipix0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t ipix = ipix0;
  const int s = blockIdx.y;
  unsigned long long mySegs = 0;
  const bool last = pixelLevel == P.maxPixelLevel;
  // Rays of the last pixel level run to the box boundary and differ in length (22 of 32 lanes were busy on average):
  // there a lane that has finished its ray takes the next pixel of its source from a queue, the warp stays full until
  // the queue is empty.  The earlier levels end at the split radius and keep the static pixel -> thread map.
  const bool refill = last && P.queue != nullptr;
  const int weight = __ldg(P.srcWeight + s);

  Cell c{0, 0, 0, 0, 0};
  double px = 0.5, py = 0.5, pz = 0.5, radius = 0., d1 = 0., d2 = 0., d3 = 0., dD = 0.;
  double prox = 0., proy = 0., proz = 1.;
  // ndot1 = float(weight) / 12.d0, then / 4.d0 per split level (exact)
  const double ndot = M(D((double)(float)weight, 12.), 1.0 / (double)(1LL << (2 * (pixelLevel - 1))));
  double* diag = P.diag + (size_t)s * kDiagStride;
  const double fnx = (double)(float)P.nx;
  int strategy = 0;      // 0 = no ray, 1 = proceed, 2 = split, 3 = boundary / dead
  int irLow = 0;
  unsigned raySeg = 0;   // index of the current segment along its ray
  long long myRayBase = 0;   // planned deposition: first slot-map entry of this thread's ray

  // start of ray `ip` of this level; false when the ray does not exist (parent not split, start outside the box, ...)
  auto init_ray = [&](int64_t ip) -> bool {
  ipix = ip;
  bool active = true;
  c = Cell{0, 0, 0, 0, 0};
  px = 0.5; py = 0.5; pz = 0.5; radius = 0.; d1 = 0.; d2 = 0.; d3 = 0.; dD = 0.;
  irLow = 0;
  raySeg = 0;
  {
    if (pixelLevel == 1) {
      c.leaf = __ldg(P.srcLeaf + s);
    } else {
      // ---- continuation after a split (equiSources.f90:3303-3378) ----
      const int64_t parent = ipix >> 2;
      const int q = (int)(ipix & 3);
      const RayState ps = P.stateIn[(size_t)s * (npix >> 2) + parent];
      if (ps.strategy != 2) active = false;
      else {
        const int pl = __ldg(P.level + ps.leaf);
        const int X = __ldg(P.leafX + ps.leaf), Y = __ldg(P.leafY + ps.leaf), Z = __ldg(P.leafZ + ps.leaf);
        double ax = ps.x, ay = ps.y, az = ps.z;  // absoluteCoordinates (:3011-3047)
        for (int t = 0; t < pl; t++) {
          ax = ((X >> t) & 1) ? A(M(0.5, ax), 0.5) : M(0.5, ax);
          ay = ((Y >> t) & 1) ? A(M(0.5, ay), 0.5) : M(0.5, ay);
          az = ((Z >> t) & 1) ? A(M(0.5, az), 0.5) : M(0.5, az);
        }
        const double xbase = D(A((double)(float)(X >> pl), ax), fnx);
        const double ybase = D(A((double)(float)(Y >> pl), ay), fnx);
        const double zbase = D(A((double)(float)(Z >> pl), az), fnx);
        const double* pd = P.pixDir + 3 * (pix_offset(pixelLevel - 1) + parent);
        const double rn = D(ps.radius, fnx);
        double xb = 0, yb = 0, zb = 0;
        // every sibling that starts outside the box is counted as boundary loss; the reference leaves
        // `strategy = boundary` set afterwards, so the LATER siblings inside the box are skipped without being counted
        bool earlierOut = false;
        for (int sib = 0; sib <= q; sib++) {
          const double* cd = P.pixDir + 3 * (pix_offset(pixelLevel) + 4 * parent + sib);
          xb = A(xbase, M(rn, S(__ldg(cd), __ldg(pd))));
          yb = A(ybase, M(rn, S(__ldg(cd + 1), __ldg(pd + 1))));
          zb = A(zbase, M(rn, S(__ldg(cd + 2), __ldg(pd + 2))));
          const bool out = xb < 0. || xb > 1. || yb < 0. || yb > 1. || zb < 0. || zb > 1.;
          if (sib < q) earlierOut = earlierOut || out;
          else if (out) {
            const double tmp = D(M(ps.radius, P.boxSize), M(fnx, P.kpc));
            for (int ir = 0; ir < 7; ir++)
              if (P.outRadius[ir] > tmp) atomicAdd(diag + 7 + ir, ndot);  // ndot1/4. of the parent
            active = false;
          }
        }
        if (earlierOut) active = false;
        if (active) {  // localizeSplitContinuationCell (:3049-3118)
          int bi = (int)M(xb, fnx), bj = (int)M(yb, fnx), bk = (int)M(zb, fnx);
          if (bi >= P.nx || bj >= P.nx || bk >= P.nx) {  // coordinate exactly 1: the reference indexes cell(nx+1)
            atomicExch(P.err, RTB200_ERR_CHECKPOINT);
            active = false;
          } else {
            double xn = S(M(xb, fnx), (double)(float)bi), yn = S(M(yb, fnx), (double)(float)bj),
                   zn = S(M(zb, fnx), (double)(float)bk);
            int node = (bi * P.nx + bj) * P.nx + bk, lvl = 0, ch;
            while ((ch = __ldg(P.child + node)) >= 0) {
              const int i = xn < 0.5 ? 0 : 1, j = yn < 0.5 ? 0 : 1, k = zn < 0.5 ? 0 : 1;
              xn = i ? S(M(2., xn), 1.) : M(2., xn);
              yn = j ? S(M(2., yn), 1.) : M(2., yn);
              zn = k ? S(M(2., zn), 1.) : M(2., zn);
              node = ch + ((i << 2) | (j << 1) | k);
              bi = 2 * bi + i; bj = 2 * bj + j; bk = 2 * bk + k;
              lvl++;
            }
            if (xn < 0. || xn > 1. || yn < 0. || yn > 1. || zn < 0. || zn > 1.) {  // checkPoint (:2962)
              atomicExch(P.err, RTB200_ERR_CHECKPOINT);
              active = false;
            }
            c.leaf = -ch - 1; c.lvl = lvl; c.X = bi; c.Y = bj; c.Z = bk;
            px = xn; py = yn; pz = zn;
            radius = ps.radius; d1 = ps.d1; d2 = ps.d2; d3 = ps.d3; dD = ps.dD;
          }
        }
      }
    }
  }

  if (active) {
    if (pixelLevel == 1) {
      c.lvl = __ldg(P.level + c.leaf);
      c.X = __ldg(P.leafX + c.leaf); c.Y = __ldg(P.leafY + c.leaf); c.Z = __ldg(P.leafZ + c.leaf);
    }
    const double* dir = P.pixDir + 3 * (pix_offset(pixelLevel) + ipix);
    prox = __ldg(dir); proy = __ldg(dir + 1); proz = __ldg(dir + 2);
  }
  return active;
  };  // init_ray

  const double* LT = P.logTab + (size_t)s * P.planes * kPlane * kSlots;
  const double rmaxL = P.rmax[pixelLevel];
  bool have0 = false;      // static map: this thread's own ray exists
  bool exhausted = weight <= 0;
  if (!refill) {
    have0 = ipix0 < npix && weight > 0 && init_ray(ipix0);
    strategy = have0 ? 1 : 0;
    if (DEPOSIT == 2 && have0) myRayBase = __ldg(P.rayBase + (long long)s * P.raysPerSource + pix_offset(pixelLevel) + ipix0);
  }
  {
    for (;;) {
      if (refill) {
        const unsigned full = 0xffffffffu;
        const bool want = strategy != 1 && !exhausted;
        const unsigned idle = __ballot_sync(full, want);
        if (idle) {
          const int lane = threadIdx.x & 31, leader = __ffs(idle) - 1;
          int base = 0;
          if (lane == leader) base = atomicAdd(P.queue + s, __popc(idle));
          base = __shfl_sync(full, base, leader);
          if (want) {
            const int64_t ip = base + __popc(idle & ((1u << lane) - 1u));
            if (ip < npix) strategy = init_ray(ip) ? 1 : 0;
            else exhausted = true;
          }
        }
        if (__ballot_sync(full, strategy == 1 || !exhausted) == 0) break;
        if (strategy != 1) continue;
      } else if (strategy != 1) {
        break;
      }
      const double oldRadius = radius;
      // ---- drawSegment (:2412-2595) ----
      const double tmp1 = proz > 0. ? D(S(1., pz), proz) : D(-pz, proz);
      const double tmp2 = prox > 0. ? D(S(1., px), prox) : D(-px, prox);
      const double tmp3 = proy > 0. ? D(S(1., py), proy) : D(-py, proy);
      int dirn; double tmp;
      if (tmp1 < fmin(tmp2, tmp3)) { dirn = 1; tmp = tmp1; }
      else if (tmp2 < fmin(tmp1, tmp3)) { dirn = 2; tmp = tmp2; }
      else { dirn = 3; tmp = tmp3; }
      const double scale = (double)(1 << c.lvl);
      const Cell here = c;
      double len;
      int face = 0;
      if (A(M(radius, scale), tmp) < rmaxL || last) {
        len = tmp;
        radius = A(radius, D(tmp, scale));
        const double ex = A(px, M(tmp, prox)), ey = A(py, M(tmp, proy)), ez = A(pz, M(tmp, proz));
        int side; bool inside;
        if (dirn == 1) {
          side = proz < 0. ? 0 : 1;
          double a = ex, b = ey;
          inside = step_neighbour(P, c, 2, side, a, b);
          if (inside) { pz = side == 0 ? 1. : 0.; px = a; py = b; }
        } else if (dirn == 2) {
          side = prox < 0. ? 0 : 1;
          double a = ey, b = ez;
          inside = step_neighbour(P, c, 0, side, a, b);
          if (inside) { px = side == 0 ? 1. : 0.; py = a; pz = b; }
        } else {
          side = proy < 0. ? 0 : 1;
          double a = ex, b = ez;
          inside = step_neighbour(P, c, 1, side, a, b);
          if (inside) { py = side == 0 ? 1. : 0.; px = a; pz = b; }
        }
        face = dirn * 2 + side;
        if (!inside) strategy = 3;
        else if (px < 0. || px > 1. || py < 0. || py > 1. || pz < 0. || pz > 1.) {
          atomicExch(P.err, RTB200_ERR_CHECKPOINT);
          strategy = 3;
          continue;
        }
      } else if (M(radius, scale) >= rmaxL) {
        strategy = 2;
        len = 0.;
      } else {
        strategy = 2;
        tmp = S(rmaxL, M(radius, scale));
        len = tmp;
        radius = A(radius, D(tmp, scale));
        px = A(px, M(tmp, prox)); py = A(py, M(tmp, proy)); pz = A(pz, M(tmp, proz));
      }
      mySegs++;
      raySeg++;
      if (TRACE) {
        const unsigned long long slot = atomicAdd(P.traceLen, 1ULL);
        if ((long long)slot < P.traceCap) {
          P.trace[2 * slot] = ((long long)here.leaf << 32) | ((long long)pixelLevel << 28) | ((long long)ipix << 8) | face;
          P.trace[2 * slot + 1] = ((long long)s << 52) | ((long long)pixelLevel << 48) | ((long long)ipix << 24) |
                                  (long long)(raySeg - 1);
        }
      }

      // ---- optical depths of the segment (:3176-3196) ----
      const int64_t lf = here.leaf;
      const double cellSize = P.cellSize[here.lvl];
      const double plen = M(cellSize, len);
      const double hi = __ldg(P.HI + lf);
      const double tau1 = M(M(plen, hi), (double)6.3e-18f);
      const double tau2 = M(M(plen, __ldg(P.HeI + lf)), (double)7.42e-18f);
      const double tau3 = M(M(plen, __ldg(P.HeII + lf)), (double)1.58e-18f);
      double tauD = 0.;
      if (P.dust == 1) tauD = D(M(M(M(plen, hi), (double)5.4116737e-22f), __ldg(P.abun2 + lf)), (double)0.2f);
      else if (P.dust == 2)
        tauD = D(M(M(D(M(M(plen, (double)0.76f), __ldg(P.rho + lf)), (double)1.6726231e-24f), (double)5.4116737e-22f),
                   __ldg(P.abun2 + lf)), (double)0.2f);

      // ---- escape diagnostics (:3198-3233) ----
      {
        // the reference tests all 7 output radii on every segment; radii only grow along a ray, so the exact test is
        // only evaluated for radii that can lie inside [oldRadius, radius] (1e-9 relative safety margin)
        while (irLow < 7 && P.outRadiusCells[irLow] * (1. + 1e-9) < oldRadius) irLow++;
        if (irLow < 7 && P.outRadiusCells[irLow] <= radius * (1. + 1e-9)) {
          const double t1 = D(M(oldRadius, P.boxSize), fnx), t2 = D(M(radius, P.boxSize), fnx);
          for (int ir = irLow; ir < 7; ir++) {
            const double tr = P.outRadiusKpc[ir];
            if (tr >= t1 && tr <= t2) {
              const double ratio = D(S(tr, t1), S(t2, t1));
              atomicAdd(diag + ir, M(ndot, exp_ref<PORTABLE>(-A(A(M(ratio, A(tau1, tauD)), d1), dD))));
              if (ir == 6) {
                const double o1 = A(M(ratio, tau1), d1), o2 = A(M(ratio, tau2), d2), o3 = A(M(ratio, tau3), d3),
                             oD = A(M(ratio, tauD), dD);
                atomicAdd(diag + 14, M(ndot, exp_ref<PORTABLE>(-oD)));
                for (int ie = 0; ie < 300; ie++) {
                  const double a1 = M(__ldg(P.outSigma + ie), o1), a2 = M(__ldg(P.outSigma + 300 + ie), o2),
                               a3 = M(__ldg(P.outSigma + 600 + ie), o3), aD = M(__ldg(P.outSigma + 900 + ie), oD);
                  atomicAdd(diag + 16 + ie, M(ndot, exp_ref<PORTABLE>(-A(A(A(a1, a2), a3), aD))));
                }
              }
            }
          }
        }
        if (strategy == 3) {
          const double tb = D(M(radius, P.boxSize), M(fnx, P.kpc));
          for (int ir = 0; ir < 7; ir++)
            if (P.outRadius[ir] > tb) atomicAdd(diag + 7 + ir, ndot);
        }
      }
      if (fmin(fmin(A(d1, tau1), A(d2, tau2)), fmin(A(d3, tau3), A(dD, tauD))) > 100.) strategy = 3;

      // ---- rates for the entire cell (:3247-3260) ----
      double dep[6];  // krate24, krate25, krate26, crate24, crate25, crate26
      if (FAITHFUL) {
        double a, b, ea, eb;
        int st = rates_faithful<PORTABLE>(P, LT, 0, d1, d2, d3, dD, a, ea);
        if (!st) st = rates_faithful<PORTABLE>(P, LT, 0, A(d1, tau1), d2, d3, dD, b, eb);
        dep[0] = M(ndot, S(a, b)); dep[3] = M(ndot, S(ea, eb));
        if (!st) st = rates_faithful<PORTABLE>(P, LT, 1, d1, d2, d3, dD, a, ea);
        if (!st) st = rates_faithful<PORTABLE>(P, LT, 1, d1, A(d2, tau2), d3, dD, b, eb);
        dep[2] = M(ndot, S(a, b)); dep[5] = M(ndot, S(ea, eb));
        if (!st) st = rates_faithful<PORTABLE>(P, LT, 2, d1, d2, d3, dD, a, ea);
        if (!st) st = rates_faithful<PORTABLE>(P, LT, 2, d1, d2, A(d3, tau3), dD, b, eb);
        dep[1] = M(ndot, S(a, b)); dep[4] = M(ndot, S(ea, eb));
        if (st) { atomicExch(P.err, st); strategy = 3; continue; }
      } else {
        const DepthIdx q = depth_index_fast(d1, d2, d3, dD, P.dust);
        if (q.status < 0) { atomicExch(P.err, RTB200_ERR_IDEPTH); strategy = 3; continue; }
        if (q.status == 1) {
          for (int i = 0; i < 6; i++) dep[i] = 0.;
        } else {
          double n, h;
          rates_fast<0>(P, LT, q, d1, tau1, sT, n, h); dep[0] = ndot * n; dep[3] = ndot * h;
          rates_fast<1>(P, LT, q, d2, tau2, sT, n, h); dep[2] = ndot * n; dep[5] = ndot * h;
          rates_fast<2>(P, LT, q, d3, tau3, sT, n, h); dep[1] = ndot * n; dep[4] = ndot * h;
        }
      }
      if (DEPOSIT == 2) {
        // planned: this segment's slot in the leaf-ordered record array is known from the plan pass
        const unsigned int slot = __ldg(P.slotMap + myRayBase + (raySeg - 1));
        double2* rec = reinterpret_cast<double2*>(P.recVal6 + (size_t)slot * 8);
        rec[0] = make_double2(dep[0], dep[1]); rec[1] = make_double2(dep[2], dep[3]); rec[2] = make_double2(dep[4], dep[5]);
        rec[3] = make_double2(0., 0.);   // the whole 64-byte record is written: no read-modify-write of a partial DRAM atom
      } else if (DEPOSIT == 1) {
        // one record per segment; slots are handed out per warp (one counter update for the active lanes)
        const unsigned m = __activemask();
        const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(P.recCount, (unsigned long long)__popc(m));
        base = __shfl_sync(m, base, leader);
        const long long slot = (long long)(base + __popc(m & ((1u << lane) - 1u)));
        if (slot < P.recCap) {
          const long long ray = (long long)s * P.raysPerSource + pix_offset(pixelLevel) + ipix;
          const long long seg = raySeg - 1 < 4095 ? (long long)(raySeg - 1) : 4095;
          // the plan needs one key per segment: a ray of more than 4095 segments cannot be planned (use mode 1)
          if (P.rayCount && raySeg - 1 >= 4095) atomicExch(P.err, RTB200_ERR_ARG);
          P.recKey[slot] = ((long long)lf << 32) | ((ray & 0xFFFFF) << 12) | seg;
#pragma unroll
          for (int i = 0; i < 6; i++) P.recVal[(size_t)i * P.recCap + slot] = dep[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 6; i++)
          if (dep[i] != 0.) atomicAdd(P.rates + (size_t)i * P.nleaf + lf, dep[i]);
      }

      d1 = A(d1, tau1); d2 = A(d2, tau2); d3 = A(d3, tau3); dD = A(dD, tauD);
    }
  }

  if (DEPOSIT == 1 && P.rayCount && !refill && ipix0 < npix)   // plan pass: segments of this ray object
    P.rayCount[(long long)s * P.raysPerSource + pix_offset(pixelLevel) + ipix0] = have0 ? raySeg : 0u;

  // warp-aggregated segment count
  for (int o = 16; o; o >>= 1) mySegs += __shfl_down_sync(0xffffffffu, mySegs, o);
  if ((threadIdx.x & 31) == 0 && mySegs) atomicAdd(P.nseg, mySegs);

  // highestPixelLevel (equiSources.f90:3316): the deepest pixel level a split of this source has opened.  Slot 15 of
  // the diagnostics holds it as a double; the bit patterns of non-negative doubles order like integers.
  if (!last && have0 && strategy == 2)
    atomicMax(reinterpret_cast<unsigned long long*>(diag + 15),
              (unsigned long long)__double_as_longlong((double)(pixelLevel + 1)));

  if (!last && ipix0 < npix) {
    RayState out;
    out.x = px; out.y = py; out.z = pz; out.radius = radius; out.d1 = d1; out.d2 = d2; out.d3 = d3; out.dD = dD;
    out.leaf = c.leaf; out.strategy = (have0 && strategy == 2) ? 2 : 3;
    out.pad[0] = out.pad[1] = 0;
    P.stateOut[(size_t)s * npix + ipix0] = out;
  }
}

// --------------------------------------------------------------------------------------------------------------------
// Table sums (stellarBetaTable.f90:217-285): thread = one (idepth1, idepth2, idepth3, idepthDust) entry, sequential
// over the 399 frequency bins in the reference's order; the six results are stored as logarithms (the lookup takes
// log of every corner it reads: equiSources.f90:4205-4238), optionally also raw.
struct TableParams {
  const double* freq;   // [7][400]: r24, r26, r25, rD, ew1, ew2, ew3   (ew_r = (nu - nu_r) * eV_to_erg)
  int thr[3];           // first frequency index with nu >= nu_r
  const double* dtmp;   // [nsrc][400] photons/s per bin
  double* logTab;       // [nsrc][planes][kPlane][kSlots]
  double* rawTab;       // optional [nsrc][6][planes][kPlane]
  int planes;
};

template <bool PORTABLE>
__global__ void __launch_bounds__(128) point_table_kernel(const __grid_constant__ TableParams T) {
  __shared__ double sh[8][kNfreq];
  const int s = blockIdx.y;
  for (int i = threadIdx.x; i < 7 * kNfreq; i += blockDim.x) sh[i / kNfreq][i % kNfreq] = T.freq[i];
  for (int i = threadIdx.x; i < kNfreq; i += blockDim.x) sh[7][i] = T.dtmp[(size_t)s * kNfreq + i];
  __syncthreads();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= T.planes * kPlane) return;
  const int i1 = e % 11, i2 = (e / 11) % 11, i3 = (e / 121) % 11, iD = e / kPlane;
  // float(idepth)/float(ndepth)*maxOpticalDepth: a single-precision quotient (stellarBetaTable.f90:237-244)
  const double g1 = M((double)__fdiv_rn((float)i1, 10.f), 10.), g2 = M((double)__fdiv_rn((float)i2, 10.f), 10.),
               g3 = M((double)__fdiv_rn((float)i3, 10.f), 10.), gD = M((double)__fdiv_rn((float)iD, 10.f), 10.);
  double R[3] = {0., 0., 0.}, E[3] = {0., 0., 0.};
  for (int i = 1; i < kNfreq; i++) {
    const double x = A(A(A(M(sh[0][i], g1), M(sh[1][i], g2)), M(sh[2][i], g3)), M(sh[3][i], gD));
    const double a = M(sh[7][i], PORTABLE ? rtb_pm::pm_exp(-x) : exp(-x));
#pragma unroll
    for (int r = 0; r < 3; r++)
      if (i >= T.thr[r]) {
        R[r] = A(R[r], a);
        E[r] = A(E[r], M(sh[4 + r][i], a));
      }
  }
  const size_t base = (size_t)s * 6 * T.planes * kPlane;                          // raw copy: the reference's layout
  double* node = T.logTab + ((size_t)s * T.planes * kPlane + e) * kSlots;         // node-major log-tables
  for (int r = 0; r < 3; r++) {
    const size_t oR = base + (size_t)r * T.planes * kPlane + e, oE = base + (size_t)(3 + r) * T.planes * kPlane + e;
    node[2 * r] = PORTABLE ? rtb_pm::pm_log(R[r]) : log(R[r]);
    node[2 * r + 1] = PORTABLE ? rtb_pm::pm_log(E[r]) : log(E[r]);
    if (T.rawTab) { T.rawTab[oR] = R[r]; T.rawTab[oE] = E[r]; }
  }
  node[6] = 0.; node[7] = 0.;
}

// Segmented, atomic-free deposition.  The march kernel emits one (leaf, 6 deposits) record per segment; the records'
// 64-bit keys (leaf, ray, segment index) are radix-sorted (cub), which makes the order of the additions to a cell a
// function of the rays alone, and one thread per run of equal leaves adds the run up and updates the cell once.
__global__ void iota_kernel(uint32_t* idx, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    idx[i] = (uint32_t)i;
}

__global__ void segmented_reduce_kernel(const long long* __restrict__ key, const uint32_t* __restrict__ idx,
                                        const double* __restrict__ val, long long n, long long cap,
                                        double* __restrict__ rates, int64_t nleaf) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int leaf = (int)(key[i] >> 32);
    if (i > 0 && (int)(key[i - 1] >> 32) == leaf) continue;  // not the head of its run
    double sum[6] = {0., 0., 0., 0., 0., 0.};
    for (long long j = i; j < n && (int)(key[j] >> 32) == leaf; j++) {
      const uint32_t r = idx[j];
#pragma unroll
      for (int f = 0; f < 6; f++) sum[f] = __dadd_rn(sum[f], val[(size_t)f * cap + r]);
    }
#pragma unroll
    for (int f = 0; f < 6; f++) {
      double* p = rates + (size_t)f * nleaf + leaf;
      *p = __dadd_rn(*p, sum[f]);
    }
  }
}

// ---- planned deposition --------------------------------------------------------------------------------------------
// The ray geometry (which leaves a ray crosses, where it splits) depends on the grid and the sources only -- not on the
// absorber densities, which change from one outer iteration to the next (without dust no ray ends early:
// equiSources.f90:3241 takes the minimum over four depths, one of which stays 0).  The first pass over a source batch
// sorts its (leaf, ray, segment) keys once and keeps, for every segment of every ray, its position in that order; later
// passes write every deposit straight to its slot and add up each leaf's contiguous run: no atomics, no sort, and the
// same fixed summation order every time.
__global__ void plan_slot_kernel(const long long* __restrict__ sortedKey, long long n, int raysPerBatchBits,
                                 const long long* __restrict__ rayBase, unsigned int* __restrict__ slotMap,
                                 int32_t* __restrict__ sortedLeaf) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
    const long long k = sortedKey[p];
    const long long ray = (k >> 12) & 0xFFFFF, seg = k & 0xFFF;
    slotMap[rayBase[ray] + seg] = (unsigned int)p;
    sortedLeaf[p] = (int32_t)(k >> 32);
  }
}

// head of a run of equal leaves in the sorted order (selection predicate of the plan pass)
struct RunHead {
  const int32_t* sortedLeaf;
  __host__ __device__ bool operator()(long long i) const { return i == 0 || sortedLeaf[i - 1] != sortedLeaf[i]; }
};

// one thread per run of equal leaves (runStart lists the heads; runStart[nruns] = n): every thread streams its own
// stretch of 64-byte records and adds them up in order -- a fixed order, no atomics; a leaf occurs in one run per batch
__global__ void planned_reduce_kernel(const int32_t* __restrict__ sortedLeaf, const long long* __restrict__ runStart,
                                      long long nruns, long long n, const double* __restrict__ val8,
                                      double* __restrict__ rates, int64_t nleaf) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < nruns; r += (long long)gridDim.x * blockDim.x) {
    const long long lo = runStart[r], hi = r + 1 < nruns ? runStart[r + 1] : n;
    const int leaf = sortedLeaf[lo];
    double sum[6] = {0., 0., 0., 0., 0., 0.};
    for (long long j = lo; j < hi; j++) {
      const double2* q = reinterpret_cast<const double2*>(val8 + (size_t)j * 8);
      const double2 a = __ldcs(q), b = __ldcs(q + 1), c = __ldcs(q + 2);     // streamed once: evict first
      sum[0] = __dadd_rn(sum[0], a.x); sum[1] = __dadd_rn(sum[1], a.y); sum[2] = __dadd_rn(sum[2], b.x);
      sum[3] = __dadd_rn(sum[3], b.y); sum[4] = __dadd_rn(sum[4], c.x); sum[5] = __dadd_rn(sum[5], c.y);
    }
#pragma unroll
    for (int f = 0; f < 6; f++) {
      double* p = rates + (size_t)f * nleaf + leaf;
      *p = __dadd_rn(*p, sum[f]);
    }
  }
}

// exclusive scan of the per-ray segment counts (a batch has at most ~1e6 ray objects: one block, sequential chunks)
__global__ void ray_base_kernel(const unsigned int* __restrict__ count, long long n, long long* __restrict__ base) {
  __shared__ long long part[1024];
  const long long per = (n + 1023) / 1024;
  const long long lo = threadIdx.x * per, hi = lo + per < n ? lo + per : n;
  long long s = 0;
  for (long long i = lo; i < hi; i++) s += count[i];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long run = 0;
    for (int t = 0; t < 1024; t++) { const long long v = part[t]; part[t] = run; run += v; }
  }
  __syncthreads();
  long long run = part[threadIdx.x];
  for (long long i = lo; i < hi; i++) { base[i] = run; run += count[i]; }
}

__global__ void gather_kernel(const double* __restrict__ src, const int32_t* __restrict__ idx, int n, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = src[idx[i]];
}

// Scratch requests of one call, served from a pool the context keeps between calls (request i reuses slot i when it
// is large enough): no cudaMalloc / cudaFree -- and no implicit device synchronisation -- in the steady state.
struct Scratch {
  std::vector<std::pair<void*, size_t>>& pool;
  size_t cursor = 0;
  explicit Scratch(std::vector<std::pair<void*, size_t>>& p) : pool(p) {}
  template <class T> int get(T** p, size_t count) {
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    if (cursor == pool.size()) pool.push_back({nullptr, 0});
    auto& sl = pool[cursor++];
    if (sl.second < bytes) {
      if (sl.first) cudaFree(sl.first);
      sl = {nullptr, 0};
      cudaError_t e = cudaMalloc(&sl.first, bytes);
      if (e != cudaSuccess) {
        set_cuda_error("cudaMalloc", e, __FILE__, __LINE__);
        sl.first = nullptr;
        *p = nullptr;
        return e == cudaErrorMemoryAllocation ? RTB200_ERR_NOMEM : RTB200_ERR_CUDA;
      }
      sl.second = bytes;
    }
    *p = (T*)sl.first;
    return RTB200_OK;
  }
};

// cached ray geometry of the planned deposition: per source batch the slot of every (ray, segment) and the leaf of
// every slot
struct PointPlan {
  std::string key;
  int batch = 0;
  struct Batch {
    long long nrec = 0;
    long long* rayBase = nullptr;
    unsigned int* slotMap = nullptr;
    int32_t* sortedLeaf = nullptr;
    long long* runStart = nullptr;   // heads of the runs of equal leaves
    long long nruns = 0;
  };
  std::vector<Batch> batches;
  void release() {
    for (auto& b : batches) { cudaFree(b.rayBase); cudaFree(b.slotMap); cudaFree(b.sortedLeaf); cudaFree(b.runStart); }
    batches.clear();
    key.clear();
  }
};

}  // namespace

void point_release(Context& c) {
  PointPlan* pl = static_cast<PointPlan*>(c.pointPlan);
  if (!pl) return;
  pl->release();
  delete pl;
  c.pointPlan = nullptr;
}

// --------------------------------------------------------------------------------------------------------------------
int point_solve(Context& c, const PointInputs& in, double* dRates, double* hDiag, int64_t* nsegOut, long long* hTrace,
                long long traceCap, long long* traceLen, double* hRawTables, cudaStream_t s) {
  if (c.nleaf == 0 || !c.dRho || !c.dAbun2) return RTB200_ERR_ARG;
  if (in.nsrc < 0 || in.maxPixelLevel < 1 || in.maxPixelLevel > kMaxPixelLevel || in.dust < 0 || in.dust > 2)
    return RTB200_ERR_ARG;
  if (in.nWave < 2 || !in.wavelength || !in.lum || !in.metallicity || !in.aDust) return RTB200_ERR_ARG;
  if (in.nsrc > 0 && (!in.srcLeaf || !in.srcWeight)) return RTB200_ERR_ARG;
  for (int i = 0; i < in.nsrc; i++)
    if (in.srcLeaf[i] < 0 || in.srcLeaf[i] >= c.nleaf) return RTB200_ERR_ARG;
  RTB_CUDA(cudaSetDevice(c.device));
  RTB_CUDA(cudaEventRecord(c.evStart, s));
  c.lastLaunches = 0;
  c.lastSweepLaunches = 0;
  if (nsegOut) *nsegOut = 0;
  if (traceLen) *traceLen = 0;
  const int nsrc = in.nsrc;
  Scratch sc(c.pointPool);
  const bool portable = c.mathMode == RTB200_MATH_FAITHFUL && c.tune.portableMath;
  const bool faithful = c.mathMode == RTB200_MATH_FAITHFUL;
  const int planes = in.dust ? 11 : 1;

  // ---- host tables ----
  PointFreq F;
  point_frequency_tables(in.aDust, F);
  std::vector<double> freq(7 * kNfreq);
  static const double thrE[3] = {(double)13.598f, (double)24.587f, (double)54.418f};
  TableParams tp{};
  for (int i = 0; i < kNfreq; i++) {
    freq[i] = F.r24[i]; freq[kNfreq + i] = F.r26[i]; freq[2 * kNfreq + i] = F.r25[i]; freq[3 * kNfreq + i] = F.rD[i];
    for (int r = 0; r < 3; r++) freq[(4 + r) * kNfreq + i] = (F.nu[i] - thrE[r]) * 1.60217646e-12;
  }
  for (int r = 0; r < 3; r++) {
    int i = 0;
    while (i < kNfreq && !(F.nu[i] >= thrE[r])) i++;
    tp.thr[r] = i;
  }
  // pixel directions of levels 1..maxPixelLevel: a constant table, computed once per context
  if (c.pointDirsLevel != in.maxPixelLevel) {
    c.pointDirs.clear();
    if (int st = point_pixel_directions(in.maxPixelLevel, c.pointDirs)) return st;
    c.pointDirsLevel = in.maxPixelLevel;
  }
  const std::vector<double>& dirs = c.pointDirs;
  std::vector<double> outSig(4 * kNenergy);
  for (int e = 0; e < kNenergy; e++) {
    outSig[e] = F.out24[e]; outSig[kNenergy + e] = F.out26[e]; outSig[2 * kNenergy + e] = F.out25[e];
    outSig[3 * kNenergy + e] = F.outD[e];
  }

  double *dFreq, *dDirs, *dOutSig;
  int32_t *dSrcLeaf, *dSrcWeight;
  unsigned long long* dCounters;  // [0] nseg, [1] trace length
  if (int st = sc.get(&dFreq, freq.size())) return st;
  if (int st = sc.get(&dDirs, dirs.size())) return st;
  if (int st = sc.get(&dOutSig, outSig.size())) return st;
  if (int st = sc.get(&dSrcLeaf, (size_t)nsrc)) return st;
  if (int st = sc.get(&dSrcWeight, (size_t)nsrc)) return st;
  if (int st = sc.get(&dCounters, 2)) return st;
  RTB_CUDA(cudaMemcpyAsync(dFreq, freq.data(), freq.size() * 8, cudaMemcpyHostToDevice, s));
  RTB_CUDA(cudaMemcpyAsync(dDirs, dirs.data(), dirs.size() * 8, cudaMemcpyHostToDevice, s));
  RTB_CUDA(cudaMemcpyAsync(dOutSig, outSig.data(), outSig.size() * 8, cudaMemcpyHostToDevice, s));
  RTB_CUDA(cudaMemsetAsync(dCounters, 0, 16, s));
  if (nsrc == 0) {
    RTB_CUDA(cudaEventRecord(c.evStop, s));
    c.statsPending = true; c.sweepTimed = false; c.lastAlgBytes = 0;
    return RTB200_OK;
  }
  RTB_CUDA(cudaMemcpyAsync(dSrcLeaf, in.srcLeaf, (size_t)nsrc * 4, cudaMemcpyHostToDevice, s));
  RTB_CUDA(cudaMemcpyAsync(dSrcWeight, in.srcWeight, (size_t)nsrc * 4, cudaMemcpyHostToDevice, s));

  // metallicity of every source's host cell (equiSources.f90:1282-1293) -> per-source photon spectrum
  double* dAb;
  if (int st = sc.get(&dAb, (size_t)nsrc)) return st;
  gather_kernel<<<(nsrc + 255) / 256, 256, 0, s>>>(c.dAbun2, dSrcLeaf, nsrc, dAb);
  std::vector<double> hAb((size_t)nsrc);
  RTB_CUDA(cudaMemcpyAsync(hAb.data(), dAb, (size_t)nsrc * 8, cudaMemcpyDeviceToHost, s));
  RTB_CUDA(cudaStreamSynchronize(s));
  // (the GPU idles while the host prepares the spectra: keep this short -- binary search for the wavelength bracket,
  // the pixel directions cached above.  Host threads made the typical call faster still but the slow calls slower.)
  std::vector<double> dtmp((size_t)nsrc * kNfreq);
  for (int i = 0; i < nsrc; i++) {
    int iMetal; double coefMetal;
    if (in.forceMetal) { iMetal = in.forceMetal; coefMetal = in.forceCoefMetal; }
    else point_metal_bracket(hAb[i], in.metallicity, &iMetal, &coefMetal);
    point_source_spectrum(F, in.nWave, in.wavelength, in.lum, in.coefSpectrum, iMetal, coefMetal, &dtmp[(size_t)i * kNfreq]);
  }

  // ---- batches of sources ----
  int64_t npixMax = 12LL << (2 * (std::max(in.maxPixelLevel, 2) - 2));  // states of level maxPixelLevel-1
  const size_t perSrc = (size_t)kSlots * planes * kPlane * 8 + 2 * (size_t)npixMax * sizeof(RayState) + kDiagStride * 8 + kNfreq * 8;
  size_t freeB = 0, totalB = 0;
  RTB_CUDA(cudaMemGetInfo(&freeB, &totalB));
  int batch = (int)std::min<size_t>((size_t)nsrc, std::max<size_t>(1, (freeB / 2) / perSrc));
  batch = std::min(batch, 16384);
  if (c.tune.pointBatch > 0) batch = std::min(batch, c.tune.pointBatch);
  const int raysPerSource = (int)(4 * ((1LL << (2 * in.maxPixelLevel)) - 1));
  // deposit mode: 0 RED.ADD, 1 records + sort every pass, 2 planned (sort once per geometry; needs rays that never end
  // early, i.e. no dust: equiSources.f90:3241)
  int depositMode = (hTrace && traceCap > 0) ? 0 : c.tune.pointDeposit;
  if (depositMode == 2 && in.dust != 0) depositMode = 1;
  // record modes: 20-bit ray index in the sort key; the batch size bounds the record buffers
  if (depositMode == 1) batch = std::max(1, std::min(std::min(batch, 16), (1 << 20) / raysPerSource));
  if (depositMode == 2) batch = std::max(1, std::min(std::min(batch, 64), (1 << 20) / raysPerSource));
  PointPlan* plan = nullptr;
  bool planning = false;
  if (depositMode == 2) {
    if (!c.pointPlan) c.pointPlan = new PointPlan();
    plan = static_cast<PointPlan*>(c.pointPlan);
    // the geometry depends on the octree, the sources that cast rays, the pixel levels and the batching
    unsigned long long h = 1469598103934665603ULL;
    for (int i = 0; i < nsrc; i++) {
      h = (h ^ (unsigned long long)(unsigned)in.srcLeaf[i]) * 1099511628211ULL;
      h = (h ^ (unsigned long long)(in.srcWeight[i] > 0)) * 1099511628211ULL;
    }
    char kb[160];
    snprintf(kb, sizeof(kb), "%p:%lld:%d:%d:%d:%d:%llx", (void*)c.tree.child, (long long)c.nleaf, c.maxLevel, in.maxPixelLevel, nsrc,
             batch, h);
    if (plan->key != kb) {
      plan->release();
      plan->key = kb;
      plan->batch = batch;
      planning = true;
    }
  }
  double *dDtmp, *dLogTab, *dRaw = nullptr, *dDiag;
  RayState *dStA, *dStB;
  if (int st = sc.get(&dDtmp, (size_t)batch * kNfreq)) return st;
  if (int st = sc.get(&dLogTab, (size_t)batch * kSlots * planes * kPlane)) return st;
  if (hRawTables) { if (int st = sc.get(&dRaw, (size_t)batch * 6 * planes * kPlane)) return st; }
  if (int st = sc.get(&dDiag, (size_t)batch * kDiagStride)) return st;
  if (int st = sc.get(&dStA, (size_t)batch * npixMax)) return st;
  if (int st = sc.get(&dStB, (size_t)batch * npixMax)) return st;
  long long* dTrace = nullptr;
  if (hTrace && traceCap > 0) { if (int st = sc.get(&dTrace, (size_t)traceCap * 2)) return st; }
  int* dQueue = nullptr;   // per-source pixel cursor of the last level (lane refill)
  if (int st = sc.get(&dQueue, (size_t)batch)) return st;
  // segmented deposition: record buffers sized from the free memory (76 B per record incl. the sort's double buffers)
  const bool segmented = depositMode == 1 || planning;      // the plan pass is a sort pass that also records the geometry
  const bool planned = depositMode == 2 && !planning;
  long long recCap = 0;
  unsigned int* dRayCount = nullptr;
  double* dRecVal6 = nullptr;
  if (planning) { if (int st = sc.get(&dRayCount, (size_t)batch * raysPerSource)) return st; }
  if (planned) {
    long long mx = 1;
    for (const auto& b : plan->batches) mx = std::max(mx, b.nrec);
    if (int st = sc.get(&dRecVal6, (size_t)mx * 8)) return st;
  }
  long long *dRecKey = nullptr, *dRecKeyOut = nullptr;
  uint32_t *dRecIdx = nullptr, *dRecIdxOut = nullptr;
  double* dRecVal = nullptr;
  void* dSortTmp = nullptr;
  size_t sortTmpBytes = 0;
  unsigned long long* dRecCount = nullptr;
  if (segmented) {
    RTB_CUDA(cudaMemGetInfo(&freeB, &totalB));
    // a leaf ray crosses at most ~2 n 2^maxLevel cells; rays of the coarser pixel levels add a few per cent
    const long long leafRays = 12LL << (2 * (in.maxPixelLevel - 1));
    const long long estimate = (long long)batch * leafRays * 2 * c.nx * (1LL << c.maxLevel) + (1 << 20);
    recCap = (long long)std::min<size_t>((size_t)(0.6 * (double)freeB) / 80, (size_t)0xFFFFFFF0u);
    recCap = std::min(recCap, estimate);
    if (c.tune.pointRecordCap > 0) recCap = std::min<long long>(recCap, c.tune.pointRecordCap);
    if (int st = sc.get(&dRecKey, (size_t)recCap)) return st;
    if (int st = sc.get(&dRecKeyOut, (size_t)recCap)) return st;
    if (int st = sc.get(&dRecIdx, (size_t)recCap)) return st;
    if (int st = sc.get(&dRecIdxOut, (size_t)recCap)) return st;
    if (int st = sc.get(&dRecVal, (size_t)6 * recCap)) return st;
    if (int st = sc.get(&dRecCount, 1)) return st;
    RTB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sortTmpBytes, dRecKey, dRecKeyOut, dRecIdx, dRecIdxOut, (long long)recCap,
                                             0, 64, s));
    char* tmp = nullptr;
    if (int st = sc.get(&tmp, sortTmpBytes)) return st;
    dSortTmp = tmp;
  }

  PointParams P{};
  P.child = c.tree.child; P.level = c.dLevel; P.leafX = c.tree.leafX; P.leafY = c.tree.leafY; P.leafZ = c.tree.leafZ;
  P.HI = c.dHI; P.HeI = c.dHeI; P.HeII = c.dHeII; P.rho = c.dRho; P.abun2 = c.dAbun2;
  P.nleaf = c.nleaf; P.nx = c.nx; P.boxSize = c.boxSize;
  P.planes = planes; P.dust = in.dust; P.maxPixelLevel = in.maxPixelLevel;
  P.pixDir = dDirs;
  double rmax[31];
  point_split_radii(rmax);
  for (int i = 0; i <= kMaxPixelLevel + 1; i++) P.rmax[i] = rmax[i];
  static const float orad[7] = {0.1f, 0.3f, 1.f, 3.f, 10.f, 30.f, 100.f};
  const double kpc = (double)1.e3f * (double)3.08568025e18f;
  for (int i = 0; i < 7; i++) { P.outRadius[i] = (double)orad[i]; P.outRadiusKpc[i] = (double)orad[i] * kpc; }
  P.kpc = kpc;
  for (int i = 0; i < 7; i++) P.outRadiusCells[i] = P.outRadiusKpc[i] * (double)c.nx / c.boxSize;
  for (int l = 0; l < 32; l++) P.cellSize[l] = l < 31 ? c.boxSize / ((double)(float)(1u << l) * (double)(float)c.nx) : 0.;
  P.outSigma = dOutSig;
  P.rates = dRates;
  P.nseg = dCounters; P.err = c.dErr;
  P.trace = dTrace; P.traceLen = dCounters + 1; P.traceCap = dTrace ? traceCap : 0;
  P.recKey = segmented ? dRecKey : nullptr; P.recVal = dRecVal; P.recCount = dRecCount; P.recCap = recCap;
  P.raysPerSource = raysPerSource;
  P.rayCount = dRayCount; P.recVal6 = dRecVal6;
  int leafBits = 1;
  while ((1LL << leafBits) < c.nleaf) leafBits++;

  for (int b0 = 0; b0 < nsrc; b0 += batch) {
    const int nb = std::min(batch, nsrc - b0);
    RTB_CUDA(cudaMemcpyAsync(dDtmp, &dtmp[(size_t)b0 * kNfreq], (size_t)nb * kNfreq * 8, cudaMemcpyHostToDevice, s));
    RTB_CUDA(cudaMemsetAsync(dDiag, 0, (size_t)nb * kDiagStride * 8, s));
    tp.freq = dFreq; tp.dtmp = dDtmp; tp.logTab = dLogTab; tp.rawTab = dRaw; tp.planes = planes;
    dim3 tg((planes * kPlane + 127) / 128, nb);
    if (portable) point_table_kernel<true><<<tg, 128, 0, s>>>(tp);
    else point_table_kernel<false><<<tg, 128, 0, s>>>(tp);
    c.lastLaunches++;
    if (hRawTables)
      RTB_CUDA(cudaMemcpyAsync(hRawTables + (size_t)b0 * 6 * planes * kPlane, dRaw, (size_t)nb * 6 * planes * kPlane * 8,
                               cudaMemcpyDeviceToHost, s));
    P.srcLeaf = dSrcLeaf + b0; P.srcWeight = dSrcWeight + b0; P.logTab = dLogTab; P.diag = dDiag;
    if (segmented) RTB_CUDA(cudaMemsetAsync(dRecCount, 0, 8, s));
    if (planning) RTB_CUDA(cudaMemsetAsync(dRayCount, 0, (size_t)nb * raysPerSource * sizeof(unsigned int), s));
    const int bi = b0 / batch;
    if (planned) {
      const PointPlan::Batch& pb = plan->batches[(size_t)bi];
      // (no clearing of the records: without dust no ray ends early, so every slot of the plan is written again)
      P.rayBase = pb.rayBase; P.slotMap = pb.slotMap;
    }
    RayState *stIn = dStA, *stOut = dStB;
    for (int L = 1; L <= in.maxPixelLevel; L++) {
      const int64_t npix = 12LL << (2 * (L - 1));
      P.stateIn = stIn; P.stateOut = stOut;
      dim3 g((unsigned)((npix + 127) / 128), nb);
      P.queue = nullptr;
      if (L == in.maxPixelLevel && c.tune.pointRefill && npix >= 1024 && depositMode != 2) {
        // last level: fewer threads than rays, every lane takes pixels from its source's queue until it is empty;
        // rays per thread ~ what keeps the device twice over-subscribed, at most 8
        const double resident = (double)c.smCount * 512.0;
        const int perThread = (int)std::max(1.0, std::min(8.0, (double)npix * nb / (2.0 * resident)));
        if (perThread > 1) {
          const int64_t threads = ((npix + perThread - 1) / perThread + 31) / 32 * 32;
          g.x = (unsigned)((threads + 127) / 128);
          RTB_CUDA(cudaMemsetAsync(dQueue, 0, (size_t)nb * sizeof(int), s));
          P.queue = dQueue;
        }
      }
      if (dTrace) {
        if (portable) point_march_kernel<true, true, true, 0><<<g, 128, 0, s>>>(P, L);
        else if (faithful) point_march_kernel<true, false, true, 0><<<g, 128, 0, s>>>(P, L);
        else point_march_kernel<false, false, true, 0><<<g, 128, 0, s>>>(P, L);
      } else if (segmented) {
        if (portable) point_march_kernel<true, true, false, 1><<<g, 128, 0, s>>>(P, L);
        else if (faithful) point_march_kernel<true, false, false, 1><<<g, 128, 0, s>>>(P, L);
        else point_march_kernel<false, false, false, 1><<<g, 128, 0, s>>>(P, L);
      } else if (planned) {
        if (portable) point_march_kernel<true, true, false, 2><<<g, 128, 0, s>>>(P, L);
        else if (faithful) point_march_kernel<true, false, false, 2><<<g, 128, 0, s>>>(P, L);
        else if (c.tune.pointMinBlocks >= 5) point_march_kernel<false, false, false, 2, 5><<<g, 128, 0, s>>>(P, L);
        else point_march_kernel<false, false, false, 2, 4><<<g, 128, 0, s>>>(P, L);
      } else {
        if (portable) point_march_kernel<true, true, false, 0><<<g, 128, 0, s>>>(P, L);
        else if (faithful) point_march_kernel<true, false, false, 0><<<g, 128, 0, s>>>(P, L);
        // FAST arithmetic, RED deposition: the register budget is a tuning knob ("point_min_blocks": 4 = 128
        // registers, 5 = 96, 6 = 80 with spills)
        else if (c.tune.pointMinBlocks >= 6) point_march_kernel<false, false, false, 0, 6><<<g, 128, 0, s>>>(P, L);
        else if (c.tune.pointMinBlocks == 5) point_march_kernel<false, false, false, 0, 5><<<g, 128, 0, s>>>(P, L);
        else point_march_kernel<false, false, false, 0, 4><<<g, 128, 0, s>>>(P, L);
      }
      c.lastLaunches++;
      c.lastSweepLaunches++;
      std::swap(stIn, stOut);
    }
    RTB_CUDA(cudaGetLastError());
    if (segmented) {
      unsigned long long nrec = 0;
      RTB_CUDA(cudaMemcpyAsync(&nrec, dRecCount, 8, cudaMemcpyDeviceToHost, s));
      RTB_CUDA(cudaStreamSynchronize(s));
      if ((long long)nrec > recCap) return RTB200_ERR_NOMEM;  // more segments than record slots: lower "point_batch"
      if (nrec > 0) {
        const int blocks = (int)std::min<long long>(((long long)nrec + 255) / 256, (long long)c.smCount * 16);
        iota_kernel<<<blocks, 256, 0, s>>>(dRecIdx, (long long)nrec);
        size_t tb = sortTmpBytes;
        RTB_CUDA(cub::DeviceRadixSort::SortPairs(dSortTmp, tb, dRecKey, dRecKeyOut, dRecIdx, dRecIdxOut, (long long)nrec, 0,
                                                 32 + leafBits, s));
        segmented_reduce_kernel<<<blocks, 256, 0, s>>>(dRecKeyOut, dRecIdxOut, dRecVal, (long long)nrec, recCap, dRates,
                                                       c.nleaf);
        c.lastLaunches += 3;
        RTB_CUDA(cudaGetLastError());
      }
      if (planning) {
        // keep the geometry: slot of every (ray, segment) in the sorted order, leaf of every slot
        PointPlan::Batch pb;
        pb.nrec = (long long)nrec;
        const size_t nrays = (size_t)nb * raysPerSource;
        RTB_CUDA(cudaMalloc((void**)&pb.rayBase, nrays * sizeof(long long)));
        RTB_CUDA(cudaMalloc((void**)&pb.slotMap, std::max<size_t>(nrec, 1) * sizeof(unsigned int)));
        RTB_CUDA(cudaMalloc((void**)&pb.sortedLeaf, std::max<size_t>(nrec, 1) * sizeof(int32_t)));
        ray_base_kernel<<<1, 1024, 0, s>>>(dRayCount, (long long)nrays, pb.rayBase);
        if (nrec > 0) {
          const int blocks = (int)std::min<long long>(((long long)nrec + 255) / 256, (long long)c.smCount * 16);
          plan_slot_kernel<<<blocks, 256, 0, s>>>(dRecKeyOut, (long long)nrec, 20, pb.rayBase, pb.slotMap, pb.sortedLeaf);
          // run heads: positions whose leaf differs from the one before (stream compaction, once per plan)
          long long* dHeads = nullptr;
          long long* dCount = nullptr;
          RTB_CUDA(cudaMalloc((void**)&dHeads, (size_t)nrec * sizeof(long long)));
          RTB_CUDA(cudaMalloc((void**)&dCount, sizeof(long long)));
          size_t tmpBytes = 0;
          cub::CountingInputIterator<long long> first(0);
          RunHead pred{pb.sortedLeaf};
          RTB_CUDA(cub::DeviceSelect::If(nullptr, tmpBytes, first, dHeads, dCount, (long long)nrec, pred, s));
          void* dTmp = nullptr;
          RTB_CUDA(cudaMalloc(&dTmp, tmpBytes ? tmpBytes : 8));
          RTB_CUDA(cub::DeviceSelect::If(dTmp, tmpBytes, first, dHeads, dCount, (long long)nrec, pred, s));
          long long nruns = 0;
          RTB_CUDA(cudaMemcpyAsync(&nruns, dCount, sizeof(long long), cudaMemcpyDeviceToHost, s));
          RTB_CUDA(cudaStreamSynchronize(s));
          RTB_CUDA(cudaMalloc((void**)&pb.runStart, (size_t)std::max<long long>(nruns, 1) * sizeof(long long)));
          RTB_CUDA(cudaMemcpyAsync(pb.runStart, dHeads, (size_t)nruns * sizeof(long long), cudaMemcpyDeviceToDevice, s));
          RTB_CUDA(cudaStreamSynchronize(s));
          pb.nruns = nruns;
          cudaFree(dHeads); cudaFree(dCount); cudaFree(dTmp);
        }
        RTB_CUDA(cudaGetLastError());
        plan->batches.push_back(pb);
        c.lastLaunches += 2;
      }
    }
    if (planned) {
      const PointPlan::Batch& pb = plan->batches[(size_t)bi];
      if (pb.nrec > 0) {
        const int blocks = (int)std::min<long long>((pb.nruns + 127) / 128, (long long)c.smCount * 32);
        planned_reduce_kernel<<<blocks, 128, 0, s>>>(pb.sortedLeaf, pb.runStart, pb.nruns, pb.nrec, dRecVal6, dRates, c.nleaf);
        c.lastLaunches += 1;
        RTB_CUDA(cudaGetLastError());
      }
    }
    if (hDiag)
      RTB_CUDA(cudaMemcpyAsync(hDiag + (size_t)b0 * kDiagStride, dDiag, (size_t)nb * kDiagStride * 8, cudaMemcpyDeviceToHost, s));
  }
  RTB_CUDA(cudaEventRecord(c.evStop, s));
  unsigned long long counters[2] = {0, 0};
  RTB_CUDA(cudaMemcpyAsync(counters, dCounters, 16, cudaMemcpyDeviceToHost, s));
  RTB_CUDA(cudaStreamSynchronize(s));
  if (nsegOut) *nsegOut = (int64_t)counters[0];
  if (dTrace) {
    const long long n = std::min<long long>((long long)counters[1], traceCap);
    RTB_CUDA(cudaMemcpy(hTrace, dTrace, (size_t)n * 16, cudaMemcpyDeviceToHost));
    if (traceLen) *traceLen = (long long)counters[1];
  }
  if (planning) {
    // the sort buffers of the plan pass are not needed again: give them back (the next call sizes its own scratch)
    RTB_CUDA(cudaStreamSynchronize(s));
    for (auto& sl : c.pointPool) cudaFree(sl.first);
    c.pointPool.clear();
  }
  c.lastAlgBytes = 136.0 * (double)counters[0];  // 5 reads + 6 read-modify-writes per segment
  c.statsPending = true;
  c.sweepTimed = false;
  int32_t err = 0;
  RTB_CUDA(cudaMemcpy(&err, c.dErr, sizeof(err), cudaMemcpyDeviceToHost));
  if (err) { cudaMemset(c.dErr, 0, 64); return err; }
  return RTB200_OK;
}

}  // namespace rtb
