// Octree build on the device: from the per-level cell lists of the driver's input grid
// (pos, lT, lnH, lx [, abun] per level; equiSources.f90:316-423) to the leaf arrays in `writeCell` order that
// rtb200_grid_set takes.  Replaces, for the data path, the driver's sequential tree construction:
//   :455-489   box from the level-1 extent, positions normalised and stored back in single precision
//   :527-578   level-1 smoothing of the second abundance on the base grid (with metals)
//   :580-618 + placeCellProjectWithVelocity :1870-1974   every level-l cell is dropped into the tree by descending l-1
//              times with `.lt.0.5` tests; children created on the way inherit tgas, rho, HI, HeI, HeII of their parent
//              and start with abun2 = 0; the target cell takes tgas = 10**lT, nH = 10**lnH, HI = nH 10**lx,
//              rho = nH mh/psi, HeI = (1-psi) rho/mhe, HeII = 0, abun2 = abun(:,2) or 0.02 (:1935-1959)
//
// Formulation.  The reference inserts cell after cell into a pointer tree; the result does not depend on that order
// except that a later duplicate overwrites an earlier one.  Here every depth d is three sorted key sets:
//   A_d  the level's own cells (packed integer coordinates at depth d; of equal keys the LAST list entry wins),
//   R_d  the nodes that are refined = ancestors at depth d of all deeper cells,
//   E_d  the nodes that exist = all base cells (d = 0) or the 8 children of every node of R_{d-1},
// and the state of a node is its own cell's (found in A_d) or, failing that, its parent's with abun2 = 0 (found in
// E_{d-1}); leaves are E_d minus R_d, and the pre-order of `writeCell` (equiSources.f90:4044-4079) is a sort by
// (base cell, octant digits).  Everything is thrust algorithms (sort, unique, vectorised binary search, transform): the
// SAME source compiles with thrust's host backend (-DTHRUST_DEVICE_SYSTEM=THRUST_DEVICE_SYSTEM_CPP), which is how the
// CPU test suite checks it bit for bit against formats.build_leaves without a GPU.
#if !defined(__CUDACC__)   // host-backend test build with a plain C++ compiler
#ifndef __host__
#define __host__
#endif
#ifndef __device__
#define __device__
#endif
#endif
#include <thrust/binary_search.h>
#include <thrust/copy.h>
#include <thrust/device_vector.h>
#include <thrust/execution_policy.h>
#include <thrust/for_each.h>
#include <thrust/gather.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/sequence.h>
#include <thrust/sort.h>
#include <thrust/transform.h>
#include <thrust/unique.h>

#include <cmath>
#include <cstdint>
#include <new>
#include <vector>

#include "../../include/rtb200.h"

#if THRUST_DEVICE_SYSTEM == THRUST_DEVICE_SYSTEM_CUDA
#include <cuda_runtime.h>
#define RTB_OCTREE_CUDA 1
#else
#define RTB_OCTREE_CUDA 0
#endif

namespace rtb_octree {

template <class T>
using DV = thrust::device_vector<T>;
typedef unsigned long long Key;   // (ix << 42) | (iy << 21) | iz at the node's own depth

constexpr int kAxisBits = 21;
constexpr Key kAxisMask = (1ULL << kAxisBits) - 1;

__host__ __device__ inline Key pack(long long ix, long long iy, long long iz) {
  return ((Key)ix << (2 * kAxisBits)) | ((Key)iy << kAxisBits) | (Key)iz;
}
__host__ __device__ inline void unpack(Key k, long long& ix, long long& iy, long long& iz) {
  ix = (long long)(k >> (2 * kAxisBits)); iy = (long long)((k >> kAxisBits) & kAxisMask); iz = (long long)(k & kAxisMask);
}
__host__ __device__ inline Key ancestor(Key k, int up) {   // coordinates halve with every level up
  long long x, y, z;
  unpack(k, x, y, z);
  return pack(x >> up, y >> up, z >> up);
}

// definitionsModule.f90: single-precision literals widened
struct Constants {
  double psi, mp, mhe;
};

// packed coordinates of an input cell at its own depth: equiSources.f90:483-489 (normalise, store as real*4), :580-618
// (int(x*nx), remainder), :1882-1932 (`.lt.0.5` descents); key = all ones marks a cell outside the box
struct CellKey {
  const float* pos;        // [ncell][3]
  double a[3], b[3];
  int nx, depth;
  __host__ __device__ Key operator()(long long i) const {
    long long c[3];
    for (int ax = 0; ax < 3; ax++) {
      const double p = (double)pos[3 * i + ax];
      const double pn = (double)(float)((p - a[ax]) / (b[ax] - a[ax]));
      const long long base = (long long)(pn * (double)nx);
      if (pn * (double)nx < 0. || base >= nx) return ~0ULL;
      double frac = pn * (double)nx - (double)base;
      long long cc = base;
      for (int t = 0; t < depth; t++) {
        const int bit = frac >= 0.5 ? 1 : 0;
        frac = 2.0 * frac - (double)bit;
        cc = 2 * cc + bit;
      }
      c[ax] = cc;
    }
    return pack(c[0], c[1], c[2]);
  }
};

struct IsLastOfRun {
  const Key* k;
  long long n;
  __host__ __device__ bool operator()(long long i) const { return i == n - 1 || k[i] != k[i + 1]; }
};

struct AncestorOf {
  int up;
  __host__ __device__ Key operator()(Key k) const { return ancestor(k, up); }
};

struct ChildOf {   // child q (0..7: x slowest, z fastest, the i, j, k loops of writeCell) of refined node r[e / 8]
  const Key* r;
  __host__ __device__ Key operator()(long long e) const {
    long long x, y, z;
    unpack(r[e >> 3], x, y, z);
    const int q = (int)(e & 7);
    return pack(2 * x + (q >> 2), 2 * y + ((q >> 1) & 1), 2 * z + (q & 1));
  }
};

// state of the nodes of one depth: six arrays over E_d
struct State {
  DV<double> tgas, rho, HI, HeI, HeII, abun2;
  void resize(size_t n) { tgas.resize(n); rho.resize(n); HI.resize(n); HeI.resize(n); HeII.resize(n); abun2.resize(n); }
};
struct StatePtr {
  double *tgas, *rho, *HI, *HeI, *HeII, *abun2;
};
inline StatePtr ptrs(State& s) {
  return StatePtr{thrust::raw_pointer_cast(s.tgas.data()), thrust::raw_pointer_cast(s.rho.data()), thrust::raw_pointer_cast(s.HI.data()),
                  thrust::raw_pointer_cast(s.HeI.data()), thrust::raw_pointer_cast(s.HeII.data()), thrust::raw_pointer_cast(s.abun2.data())};
}

// own cell (slot in A_d, or -1) or parent (slot in E_{d-1}) -> state of node e of E_d
struct FillState {
  const long long* ownSlot;     // [|E_d|] position in the level's winner list or -1
  const long long* winner;      // [|A_d|] index of the winning list entry
  const long long* parentSlot;  // [|E_d|] position of the parent in E_{d-1} (depth > 0)
  const float *lT, *lnH, *lx, *abun;   // the level's lists; abun = smoothed / raw second abundance or NULL
  StatePtr parent, out;
  Constants c;
  int depth;
  __host__ __device__ void operator()(long long e) const {
    const long long s = ownSlot[e];
    if (s >= 0) {
      const long long i = winner[s];
      const double nH = pow(10.0, (double)lnH[i]);               // 10.**lnH, single-precision argument widened
      const double rho = nH * c.mp / c.psi;
      out.tgas[e] = pow(10.0, (double)lT[i]);
      out.rho[e] = rho;
      out.HI[e] = nH * pow(10.0, (double)lx[i]);
      out.HeI[e] = (1.0 - c.psi) * rho / c.mhe * 1.0;
      out.HeII[e] = 0.;
      out.abun2[e] = abun ? (double)abun[i] : (double)0.02f;
    } else if (depth == 0) {
      out.tgas[e] = 0.; out.rho[e] = 0.; out.HI[e] = 0.; out.HeI[e] = 0.; out.HeII[e] = 0.; out.abun2[e] = 0.;
    } else {
      const long long p = parentSlot[e];
      out.tgas[e] = parent.tgas[p]; out.rho[e] = parent.rho[p]; out.HI[e] = parent.HI[p];
      out.HeI[e] = parent.HeI[p]; out.HeII[e] = parent.HeII[p];
      out.abun2[e] = 0.;                                          // :1907 children start with abun2 = 0
    }
  }
};

// position of q in the sorted array s (or -1)
struct FindIn {
  const Key* s;
  long long n;
  __host__ __device__ long long operator()(Key q) const {
    long long lo = 0, hi = n;
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (s[mid] < q) lo = mid + 1; else hi = mid;
    }
    return (lo < n && s[lo] == q) ? lo : -1;
  }
};

struct ParentKey {
  __host__ __device__ Key operator()(Key k) const { return ancestor(k, 1); }
};

// level-1 smoothing (equiSources.f90:527-578): one pass of the (1/4, 1/2, 1/4) filter along one axis of the n^3 grid,
// terms added in the reference's order: 0.25 u(i-1), then 0.5 u(i), then 0.25 u(i+1); nothing crosses the box faces
struct SmoothAxis {
  const double* u;
  int nx, axis;
  __host__ __device__ double operator()(long long c) const {
    const long long stride = axis == 0 ? (long long)nx * nx : (axis == 1 ? nx : 1);
    const long long idx = axis == 0 ? c / ((long long)nx * nx) : (axis == 1 ? (c / nx) % nx : c % nx);
    double t = 0.;
    if (idx > 0) t = t + 0.25 * u[c - stride];
    t = t + 0.5 * u[c];
    if (idx < nx - 1) t = t + 0.25 * u[c + stride];
    return t;
  }
};

struct PreorderKey {   // (base cell, octant digits) padded to the deepest level: the order of writeCell
  const Key* k;
  int nx, depth, lmax;
  __host__ __device__ Key operator()(long long i) const {
    long long x, y, z;
    unpack(k[i], x, y, z);
    Key key = (Key)(((x >> depth) * nx + (y >> depth)) * nx + (z >> depth));
    for (int d = depth - 1; d >= 0; d--) key = (key << 3) | (Key)((((x >> d) & 1) << 2) | (((y >> d) & 1) << 1) | ((z >> d) & 1));
    return key << (3 * (lmax - depth));
  }
};

struct BaseKey {
  int nx;
  __host__ __device__ Key operator()(long long c) const { return pack(c / ((long long)nx * nx), (c / nx) % nx, c % nx); }
};
struct Scatter {
  const long long* own; const long long* win; const float* ab;
  __host__ __device__ double operator()(long long e) const { return own[e] >= 0 ? (double)ab[win[own[e]]] : 0.; }
};
struct Gather {
  const double* u; const long long* slot;
  __host__ __device__ float operator()(long long i) const { return (float)u[slot[i]]; }
};

struct NotIn {   // stencil for leaves: node not found in the refined set
  const long long* found;
  __host__ __device__ bool operator()(long long i) const { return found[i] < 0; }
};

struct Result {
  int nx = 0;
  double box = 0;
  std::vector<int8_t> level;
  std::vector<double> f[6];   // HI, HeI, HeII, rho, abun2, tgas
};

template <class T>
inline const T* raw(const DV<T>& v) { return thrust::raw_pointer_cast(v.data()); }
template <class T>
inline T* raw(DV<T>& v) { return thrust::raw_pointer_cast(v.data()); }

int build(int nlevels, const int64_t* ncell, const float* const* pos, const float* const* lT, const float* const* lnH,
          const float* const* lx, const float* const* abun2, Result& R) {
  if (nlevels < 1 || nlevels > 20 || !ncell || !pos || !lT || !lnH || !lx) return RTB200_ERR_ARG;
  const int64_t n1 = ncell[0];
  int nx = (int)llround(cbrt((double)n1));
  if ((int64_t)nx * nx * nx != n1 || nx < 1) return RTB200_ERR_LEVELS;     // 'base grid needs to be of size n^3' (:427-436)
  if (((long long)nx << (nlevels - 1)) >= (1LL << kAxisBits)) return RTB200_ERR_ARG;
  // box: level-1 min / max stretched by n/(n-1) (:455-477); a handful of host flops
  double a[3], b[3];
  for (int ax = 0; ax < 3; ax++) {
    double lo = 1e300, hi = -1e300;
    for (int64_t i = 0; i < n1; i++) {
      const double p = (double)pos[0][3 * i + ax];
      lo = p < lo ? p : lo; hi = p > hi ? p : hi;
    }
    const double mid = 0.5 * (lo + hi);
    const double half = nx > 1 ? 0.5 * (hi - lo) * (double)(float)nx / (double)(float)(nx - 1) : 0.5 * (hi - lo);
    a[ax] = mid - half; b[ax] = mid + half;
  }
  const double kpc = (double)1.e3f * (double)3.08568025e18f;
  R.nx = nx;
  R.box = fabs(a[0] - b[0]) * kpc;
  Constants C;
  C.psi = (double)0.76f; C.mp = (double)1.6726231e-24f;
  C.mhe = 2.0 * ((double)1.6726231e-24f + (double)1.67492728e-24f);
  const int lmax = nlevels - 1;

  // ---- keys of every level's cells at their own depth ----
  std::vector<DV<float>> dPos(nlevels), dLT(nlevels), dLnH(nlevels), dLx(nlevels), dAb(nlevels);
  std::vector<DV<Key>> cellKey(nlevels);
  for (int d = 0; d < nlevels; d++) {
    const size_t n = (size_t)ncell[d];
    dPos[d].assign(pos[d], pos[d] + 3 * n);
    dLT[d].assign(lT[d], lT[d] + n); dLnH[d].assign(lnH[d], lnH[d] + n); dLx[d].assign(lx[d], lx[d] + n);
    if (abun2) dAb[d].assign(abun2[d], abun2[d] + n);
    cellKey[d].resize(n);
    CellKey ck;
    ck.pos = raw(dPos[d]); ck.nx = nx; ck.depth = d;
    for (int ax = 0; ax < 3; ax++) { ck.a[ax] = a[ax]; ck.b[ax] = b[ax]; }
    thrust::transform(thrust::counting_iterator<long long>(0), thrust::counting_iterator<long long>((long long)n),
                      cellKey[d].begin(), ck);
    if (n && thrust::count(cellKey[d].begin(), cellKey[d].end(), (Key)~0ULL) > 0) return RTB200_ERR_ARG;   // cell outside the box
  }

  std::vector<DV<Key>> E(nlevels), Rf(nlevels);
  std::vector<State> V(nlevels);
  DV<Key> allKeys;
  DV<int8_t> allLevel;
  State allState;
  for (int d = 0; d < nlevels; d++) {
    const long long n = (long long)ncell[d];
    // A_d: sorted unique keys of the level's cells, the last list entry of equal keys wins
    DV<Key> ak = cellKey[d];
    DV<long long> ai((size_t)n);
    thrust::sequence(ai.begin(), ai.end());
    thrust::stable_sort_by_key(ak.begin(), ak.end(), ai.begin());
    DV<Key> aKey((size_t)n);
    DV<long long> aWin((size_t)n);
    IsLastOfRun last{raw(ak), n};
    auto endK = thrust::copy_if(ak.begin(), ak.end(), thrust::counting_iterator<long long>(0), aKey.begin(), last);
    thrust::copy_if(ai.begin(), ai.end(), thrust::counting_iterator<long long>(0), aWin.begin(), last);
    const long long na = (long long)(endK - aKey.begin());
    aKey.resize((size_t)na); aWin.resize((size_t)na);
    // R_d: ancestors at depth d of all deeper cells
    size_t deeper = 0;
    for (int l = d + 1; l < nlevels; l++) deeper += (size_t)ncell[l];
    Rf[d].resize(deeper);
    size_t o = 0;
    for (int l = d + 1; l < nlevels; l++) {
      thrust::transform(cellKey[l].begin(), cellKey[l].end(), Rf[d].begin() + o, AncestorOf{l - d});
      o += (size_t)ncell[l];
    }
    thrust::sort(Rf[d].begin(), Rf[d].end());
    Rf[d].erase(thrust::unique(Rf[d].begin(), Rf[d].end()), Rf[d].end());
    // E_d: all base cells, or the children of R_{d-1}
    if (d == 0) {
      E[0].resize((size_t)n1);
      // base keys in (ix, iy, iz) order are already sorted
      thrust::transform(thrust::counting_iterator<long long>(0), thrust::counting_iterator<long long>((long long)n1), E[0].begin(),
                        BaseKey{nx});
    } else {
      E[d].resize(8 * Rf[d - 1].size());
      thrust::transform(thrust::counting_iterator<long long>(0), thrust::counting_iterator<long long>((long long)E[d].size()),
                        E[d].begin(), ChildOf{raw(Rf[d - 1])});
      thrust::sort(E[d].begin(), E[d].end());
    }
    const long long ne = (long long)E[d].size();
    // every cell of the level must exist in E_d (its ancestors are in R by construction; a level-1 cell always)
    // own slot / parent slot of every node
    DV<long long> ownSlot((size_t)ne), parentSlot((size_t)ne);
    thrust::transform(E[d].begin(), E[d].end(), ownSlot.begin(), FindIn{raw(aKey), na});
    if (d > 0) {
      DV<Key> pk((size_t)ne);
      thrust::transform(E[d].begin(), E[d].end(), pk.begin(), ParentKey());
      thrust::transform(pk.begin(), pk.end(), parentSlot.begin(), FindIn{raw(E[d - 1]), (long long)E[d - 1].size()});
    }
    // second abundance of the level's list; level 1 with metals: smoothed on the base grid first
    DV<float> abunUse;
    const float* abunPtr = nullptr;
    if (abun2) {
      abunUse = dAb[d];
      if (d == 0) {
        DV<double> u((size_t)n1), t((size_t)n1);
        // scatter (a later duplicate overwrites): base cell e takes its winner's value, cells without one stay 0
        thrust::transform(thrust::counting_iterator<long long>(0), thrust::counting_iterator<long long>((long long)n1), u.begin(),
                          Scatter{raw(ownSlot), raw(aWin), raw(dAb[0])});
        for (int rep = 0; rep < 2; rep++)
          for (int ax = 0; ax < 3; ax++) {
            thrust::transform(thrust::counting_iterator<long long>(0), thrust::counting_iterator<long long>((long long)n1), t.begin(),
                              SmoothAxis{raw(u), nx, ax});
            u.swap(t);
          }
        // gather back into the real*4 list: every list entry reads its base cell
        DV<long long> cellSlot((size_t)n);
        thrust::transform(cellKey[0].begin(), cellKey[0].end(), cellSlot.begin(), FindIn{raw(E[0]), (long long)n1});
        thrust::transform(thrust::counting_iterator<long long>(0), thrust::counting_iterator<long long>(n), abunUse.begin(),
                          Gather{raw(u), raw(cellSlot)});
      }
      abunPtr = raw(abunUse);
    }
    V[d].resize((size_t)ne);
    FillState fs;
    fs.ownSlot = raw(ownSlot); fs.winner = raw(aWin); fs.parentSlot = raw(parentSlot);
    fs.lT = raw(dLT[d]); fs.lnH = raw(dLnH[d]); fs.lx = raw(dLx[d]); fs.abun = abunPtr;
    fs.parent = d > 0 ? ptrs(V[d - 1]) : StatePtr{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    fs.out = ptrs(V[d]); fs.c = C; fs.depth = d;
    thrust::for_each(thrust::counting_iterator<long long>(0), thrust::counting_iterator<long long>(ne), fs);
    // leaves of this depth: E_d without R_d
    DV<long long> inR((size_t)ne);
    thrust::transform(E[d].begin(), E[d].end(), inR.begin(), FindIn{raw(Rf[d]), (long long)Rf[d].size()});
    DV<long long> leafIdx((size_t)ne);
    auto endL = thrust::copy_if(thrust::counting_iterator<long long>(0), thrust::counting_iterator<long long>(ne), leafIdx.begin(),
                                NotIn{raw(inR)});
    const size_t nl = (size_t)(endL - leafIdx.begin());
    leafIdx.resize(nl);
    const size_t base = allKeys.size();
    allKeys.resize(base + nl); allLevel.resize(base + nl); allState.resize(base + nl);
    DV<Key> pre((size_t)ne);
    thrust::transform(thrust::counting_iterator<long long>(0), thrust::counting_iterator<long long>(ne), pre.begin(),
                      PreorderKey{raw(E[d]), nx, d, lmax});
    thrust::gather(leafIdx.begin(), leafIdx.end(), pre.begin(), allKeys.begin() + base);
    thrust::fill(allLevel.begin() + base, allLevel.end(), (int8_t)d);
    thrust::gather(leafIdx.begin(), leafIdx.end(), V[d].tgas.begin(), allState.tgas.begin() + base);
    thrust::gather(leafIdx.begin(), leafIdx.end(), V[d].rho.begin(), allState.rho.begin() + base);
    thrust::gather(leafIdx.begin(), leafIdx.end(), V[d].HI.begin(), allState.HI.begin() + base);
    thrust::gather(leafIdx.begin(), leafIdx.end(), V[d].HeI.begin(), allState.HeI.begin() + base);
    thrust::gather(leafIdx.begin(), leafIdx.end(), V[d].HeII.begin(), allState.HeII.begin() + base);
    thrust::gather(leafIdx.begin(), leafIdx.end(), V[d].abun2.begin(), allState.abun2.begin() + base);
  }
  // ---- pre-order: sort all leaves by (base cell, octant digits) ----
  const size_t N = allKeys.size();
  DV<long long> order(N);
  thrust::sequence(order.begin(), order.end());
  thrust::sort_by_key(allKeys.begin(), allKeys.end(), order.begin());
  auto fetch = [&](DV<double>& src, std::vector<double>& dst) {
    DV<double> tmp(N);
    thrust::gather(order.begin(), order.end(), src.begin(), tmp.begin());
    dst.resize(N);
    thrust::copy(tmp.begin(), tmp.end(), dst.begin());
  };
  fetch(allState.HI, R.f[0]); fetch(allState.HeI, R.f[1]); fetch(allState.HeII, R.f[2]); fetch(allState.rho, R.f[3]);
  fetch(allState.abun2, R.f[4]); fetch(allState.tgas, R.f[5]);
  {
    DV<int8_t> tmp(N);
    thrust::gather(order.begin(), order.end(), allLevel.begin(), tmp.begin());
    R.level.resize(N);
    thrust::copy(tmp.begin(), tmp.end(), R.level.begin());
  }
  return RTB200_OK;
}

}  // namespace rtb_octree

extern "C" {

#if RTB_OCTREE_CUDA
#define RTB_OCTREE_NAME(x) rtb200_octree_##x
#else
#define RTB_OCTREE_NAME(x) rtb200_hostcheck_octree_##x   /* thrust host backend: test build of the same source */
#endif

int RTB_OCTREE_NAME(build)(int device, int nlevels, const int64_t* ncell, const float* const* pos, const float* const* lT,
                           const float* const* lnH, const float* const* lx, const float* const* abun2, int64_t* nleaf,
                           int32_t* nx, double* physicalBoxSize, void** handle) {
  if (!handle || !nleaf || !nx || !physicalBoxSize) return RTB200_ERR_ARG;
  *handle = nullptr;
#if RTB_OCTREE_CUDA
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return RTB200_ERR_CUDA;   // no CPU fallback in the product
  if (device < 0 || device >= ndev) return RTB200_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return RTB200_ERR_CUDA;
#else
  (void)device;
#endif
  rtb_octree::Result* R = new (std::nothrow) rtb_octree::Result();
  if (!R) return RTB200_ERR_NOMEM;
  int st;
  try {
    st = rtb_octree::build(nlevels, ncell, pos, lT, lnH, lx, abun2, *R);
  } catch (const std::bad_alloc&) {
    st = RTB200_ERR_NOMEM;
  } catch (...) {
    st = RTB200_ERR_CUDA;
  }
  if (st) { delete R; return st; }
  *nleaf = (int64_t)R->level.size();
  *nx = R->nx;
  *physicalBoxSize = R->box;
  *handle = R;
  return RTB200_OK;
}

int RTB_OCTREE_NAME(get)(void* handle, int8_t* level, double* HI, double* HeI, double* HeII, double* rho, double* abun2,
                         double* tgas) {
  if (!handle) return RTB200_ERR_ARG;
  rtb_octree::Result* R = static_cast<rtb_octree::Result*>(handle);
  const size_t N = R->level.size();
  if (level) std::copy(R->level.begin(), R->level.end(), level);
  double* out[6] = {HI, HeI, HeII, rho, abun2, tgas};
  for (int f = 0; f < 6; f++)
    if (out[f]) std::copy(R->f[f].begin(), R->f[f].begin() + N, out[f]);
  return RTB200_OK;
}

int RTB_OCTREE_NAME(free)(void* handle) {
  delete static_cast<rtb_octree::Result*>(handle);
  return RTB200_OK;
}

}  // extern "C"
