// Diffuse sweep on a uniform (single-level) grid: replaces the direction loop of equiSources.f90:1389-1806 for
// grids without refinement (configs 2 and 4 of BASELINE.json).
//
// Formulation (DESIGN.md "uniform sweep"):
//  * All directions of one zone share the index rotation (rotateIndicesModule.f90), so they are swept TOGETHER,
//    layer by layer along the zone's sweep axis: one thread owns one cell of the layer, reads kappa1..3 once,
//    loops over the zone's directions and adds all their contributions to J in registers -> one J update per
//    cell per zone instead of one per direction.
//  * Within a layer every cell carries the same 1..3 segments (its "pattern"); a characteristic that leaves a
//    cell sideways continues in the k+1 and/or j+1 neighbour of the SAME layer.  Along the lane axis that hand-over
//    is a warp shuffle of the neighbour lane's own result (a warp covers 31 cells + 1 recomputed halo cell); along
//    the row axis the thread recomputes what the (b-1) cell emits from the previous layer's plane value and that
//    cell's kappa.  No shared-memory exchange, no barrier in the loop, no inter-block communication.
//  * The only inter-layer state is the intensity leaving each cell through its top face: one padded plane of
//    3 doubles per cell per direction ([row][column][group]), ping-ponged in global memory.
//  * One launch per layer covers every zone task of the batch (gridDim.z); each task owns a private J accumulator
//    ("slot"), so no atomics and a fixed summation order.  Final kernels sum the slot accumulators into J.
//  * Zones that sweep along the contiguous axis of the leaf order read a z-major copy of kappa and accumulate in
//    that layout (transpose_kappa_kernel / merge_transposed_kernel) so that their lanes stay coalesced.
//  * sweep_persistent_kernel runs the whole layer loop in one launch (tiles ordered by progress words); off by default.

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "rtb200_internal.h"
#include "segment_math.cuh"

namespace rtb {

__global__ void compute_opacities_kernel(const double* __restrict__ HI, const double* __restrict__ HeI,
                                         const double* __restrict__ HeII, double* __restrict__ kappa, int64_t n,
                                         double b0, double b3, double b4, double b6, double b7, double b8) {
  // equiSources.f90:4974-4977, same association order, no FMA contraction
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double h = HI[i], he1 = HeI[i], he2 = HeII[i];
    kappa[i] = __dmul_rn(h, b0);
    kappa[n + i] = __dadd_rn(__dmul_rn(h, b3), __dmul_rn(he1, b4));
    kappa[2 * n + i] = __dadd_rn(__dadd_rn(__dmul_rn(h, b6), __dmul_rn(he1, b7)), __dmul_rn(he2, b8));
  }
}

int launch_compute_opacities(Context& c, const double* beta, cudaStream_t s) {
  int blocks = std::min<int64_t>((c.nleaf + 255) / 256, (int64_t)c.smCount * 16);
  compute_opacities_kernel<<<blocks, 256, 0, s>>>(c.dHI, c.dHeI, c.dHeII, c.dKappa, c.nleaf, beta[0], beta[3], beta[4],
                                                  beta[6], beta[7], beta[8]);
  RTB_CUDA(cudaGetLastError());
  return RTB200_OK;
}

// ---------------------------------------------------------------------------------------------------------
// one layer step of one zone
//
// thread = one cell of the layer; it loops over the zone's directions, keeping kappa (own cell and the three
// upstream neighbours of the layer) and the J sum in registers.  A segment fed by a same-layer neighbour needs that
// neighbour's outgoing intensity, which is a pure function of the previous layer's top-exit plane and of kappa:
// the thread RECOMPUTES it (one or two extra exponentials, bit-identical to what the neighbour's own thread
// computes) instead of waiting for it.  Threads are independent: no shared memory, no halo, no barrier.
// The per-step tables travel as a __grid_constant__ kernel parameter, i.e. they are read from the constant bank.
// ---------------------------------------------------------------------------------------------------------
struct StepParams {
  LayerSeg P[kMaxDirPerTask];
  const double* planeIn;   // this task's planes of the previous layer  [ndir][n+1][n+1][3 groups] (padded, see below)
  double* planeOut;
  double* acc;             // slot accumulator [3][N]
  const double* kappa;     // [3][N] in the layout the task's strides refer to (leaf order or z-major)
  int32_t origin;          // leaf index of rotated (step, 0, 0)
  int32_t sj, sk;
  int32_t ndir, laneIsK, firstInSlot, n;
};

// Planes carry one extra row and column on the upstream side (index -1) that permanently hold the boundary
// intensity, and the plane read by the first layer is filled with it: together with kappa = 0 for out-of-domain
// neighbours (exp(-0) = 1 exactly) this makes "no neighbour -> uvb" (transportRoutinesModule.f90:594-597) fall out
// of the same arithmetic as an interior cell, with no per-direction edge test.
__global__ void fill_planes_kernel(double* __restrict__ planes, int64_t perGroup, int64_t total, double u0, double u1,
                                   double u2) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int g = (int)(i % 3);  // the three frequency groups of a cell are adjacent: [direction][row][column][group]
    planes[i] = g == 0 ? u0 : (g == 1 ? u1 : u2);
  }
}

// One direction of one cell, straight-line for a given segment count and chain order.
//   NSEG: 1..3 segments.  SECL: the second segment is fed by the neighbour along the LANE axis (cell a-1, same row)
//   and the third by the neighbour along the ROW axis (cell b-1, same lane); !SECL: the other way round.
// Lane-axis hand-over = warp shuffle of the neighbour lane's own result; row-axis hand-over = recompute of what the
// (b-1) cell emits from the previous layer's plane value `upR` and its kappa `kR` (bit-identical to that cell's own
// arithmetic).  Every lane of the warp must call this (shuffles), including the halo lane and out-of-domain lanes.
//
// FAST arithmetic (segment_math.cuh: segment_fast): A[g] collects Iin (1 - e^-tau) cs over all segments and
// directions of the layer; the caller multiplies by 2^-200 / kappa once per layer.
template <int EXPV, int NSEG, bool SECL, bool GUARD>
__device__ __forceinline__ void direction_fast(const LayerSeg& P, const double (&cur)[3], const double (&upR)[3],
                                               const double (&kap)[3], const double (&kR)[3], double (&I)[3],
                                               double (&A)[3], const double* __restrict__ T) {
  const unsigned full = 0xffffffffu;
#pragma unroll
  for (int g = 0; g < 3; g++) {
    const double I1 = segment_fast<EXPV, GUARD>(cur[g], kap[g] * P.d[0], P.cs[0], T, A[g]);
    I[g] = I1;
    if (NSEG >= 2) {
      double rup = 0.;  // xy-segment output of the (b-1) cell
      if (!SECL || NSEG == 3) rup = attenuate_fast<EXPV, GUARD>(upR[g], kR[g] * P.d[0], T);
      const double in2 = SECL ? __shfl_up_sync(full, I1, 1) : rup;
      const double I2 = segment_fast<EXPV, GUARD>(in2, kap[g] * P.d[1], P.cs[1], T, A[g]);
      I[g] = I2;
      if (NSEG == 3) {
        double in3;
        if (SECL) {
          // third segment from the (b-1) cell's SECOND segment, which was fed by the (b-1, a-1) cell's xy segment
          const double x = __shfl_up_sync(full, rup, 1);
          in3 = attenuate_fast<EXPV, GUARD>(x, kR[g] * P.d[1], T);
        } else {
          in3 = __shfl_up_sync(full, I2, 1);  // the (a-1) cell's second segment
        }
        I[g] = segment_fast<EXPV, GUARD>(in3, kap[g] * P.d[2], P.cs[2], T, A[g]);
      }
    }
  }
}

// The reference's own operation sequence (RTB200_MATH_FAITHFUL, and the thin layers of FAST mode): adds
// (sum of the segments' J) / nseg * weight to acc (transportRoutinesModule.f90:953-955).
template <int NSEG, bool SECL>
__device__ __forceinline__ void direction_faithful(const LayerSeg& P, const double (&cur)[3], const double (&upR)[3],
                                                   const double (&kap)[3], const double (&kR)[3], double (&I)[3],
                                                   double (&acc)[3]) {
  const unsigned full = 0xffffffffu;
#pragma unroll
  for (int g = 0; g < 3; g++) {
    SegResult r1 = segment_update<true, 0>(cur[g], kap[g], P.d[0], 0., nullptr);
    double Jsum = r1.J;
    I[g] = r1.Iout;
    if (NSEG >= 2) {
      double rup = 0.;
      if (!SECL || NSEG == 3) rup = __dmul_rn(upR[g], exp(-__dmul_rn(kR[g], P.d[0])));
      const double in2 = SECL ? __shfl_up_sync(full, r1.Iout, 1) : rup;
      SegResult r2 = segment_update<true, 0>(in2, kap[g], P.d[1], 0., nullptr);
      I[g] = r2.Iout;
      double J3 = 0.;
      if (NSEG == 3) {
        double in3;
        if (SECL) {
          const double x = __shfl_up_sync(full, rup, 1);
          in3 = __dmul_rn(x, exp(-__dmul_rn(kR[g], P.d[1])));
        } else {
          in3 = __shfl_up_sync(full, r2.Iout, 1);
        }
        SegResult r3 = segment_update<true, 0>(in3, kap[g], P.d[2], 0., nullptr);
        I[g] = r3.Iout;
        J3 = r3.J;
      }
      // the reference sums the segments in the order xy, xz, yz (transportRoutinesModule.f90:698,818,941);
      // P.kind <= 2 <=> the second segment is the yz ray
      const bool yzSecond = P.kind <= 2;
      const double Jxz = yzSecond ? J3 : r2.J, Jyz = yzSecond ? r2.J : J3;
      Jsum = r1.J;
      if (NSEG == 3 || !yzSecond) Jsum = __dadd_rn(Jsum, Jxz);
      if (NSEG == 3 || yzSecond) Jsum = __dadd_rn(Jsum, Jyz);
    }
    acc[g] = __dadd_rn(acc[g], __dmul_rn(__ddiv_rn(Jsum, (double)NSEG), P.w));
  }
}

// dispatch on the layer's chain kind (warp-uniform).  kmax = the largest opacity this thread multiplies a path with.
// If kmax * (longest segment of the layer) <= 64 for every lane of the warp, the overflow guards of segment_fast are
// dropped (4 fewer instructions per segment update); `thick` warps take the guarded instantiation.
template <int EXPV, bool GUARD>
__device__ __forceinline__ void direction_kinds_fast(const LayerSeg& P, bool secL, const double (&cur)[3],
                                                     const double (&upR)[3], const double (&kap)[3],
                                                     const double (&kR)[3], double (&I)[3], double (&A)[3],
                                                     const double* __restrict__ T) {
  const int kind = P.kind;
  if (kind == 0) direction_fast<EXPV, 1, true, GUARD>(P, cur, upR, kap, kR, I, A, T);
  else if (kind == 1 || kind == 3) {
    if (secL) direction_fast<EXPV, 2, true, GUARD>(P, cur, upR, kap, kR, I, A, T);
    else direction_fast<EXPV, 2, false, GUARD>(P, cur, upR, kap, kR, I, A, T);
  } else {
    if (secL) direction_fast<EXPV, 3, true, GUARD>(P, cur, upR, kap, kR, I, A, T);
    else direction_fast<EXPV, 3, false, GUARD>(P, cur, upR, kap, kR, I, A, T);
  }
}

template <int EXPV>
__device__ __forceinline__ void direction_dispatch_fast(const LayerSeg& P, bool secL, double kmax, const double (&cur)[3],
                                                        const double (&upR)[3], const double (&kap)[3],
                                                        const double (&kR)[3], double (&I)[3], double (&A)[3],
                                                        const double* __restrict__ T) {
  if (__any_sync(0xffffffffu, kmax * P.dmax > 64.)) direction_kinds_fast<EXPV, true>(P, secL, cur, upR, kap, kR, I, A, T);
  else direction_kinds_fast<EXPV, false>(P, secL, cur, upR, kap, kR, I, A, T);
}

// Out of line and by value: the rare faithful branch must not force the fast path's per-direction arrays into
// local memory (a by-reference call would).
struct Tri {
  double v[3];
};
struct FaithOut {
  Tri I, acc;
};
#ifndef RTB_FAITHFUL_INLINE
#define RTB_FAITHFUL_CALL __noinline__
#else
#define RTB_FAITHFUL_CALL __forceinline__
#endif
__device__ RTB_FAITHFUL_CALL FaithOut direction_call_faithful(double d0, double d1, double d2, double w, int kind,
                                                              bool secL, Tri cur_, Tri upR_, Tri kap_, Tri kR_, Tri acc_) {
  LayerSeg P;
  P.d[0] = d0; P.d[1] = d1; P.d[2] = d2; P.w = w; P.kind = kind;
  double cur[3] = {cur_.v[0], cur_.v[1], cur_.v[2]}, upR[3] = {upR_.v[0], upR_.v[1], upR_.v[2]};
  double kap[3] = {kap_.v[0], kap_.v[1], kap_.v[2]}, kR[3] = {kR_.v[0], kR_.v[1], kR_.v[2]};
  double acc[3] = {acc_.v[0], acc_.v[1], acc_.v[2]}, I[3] = {0., 0., 0.};
  if (kind == 0) direction_faithful<1, true>(P, cur, upR, kap, kR, I, acc);
  else if (kind == 1 || kind == 3) {
    if (secL) direction_faithful<2, true>(P, cur, upR, kap, kR, I, acc);
    else direction_faithful<2, false>(P, cur, upR, kap, kR, I, acc);
  } else {
    if (secL) direction_faithful<3, true>(P, cur, upR, kap, kR, I, acc);
    else direction_faithful<3, false>(P, cur, upR, kap, kR, I, acc);
  }
  FaithOut o;
#pragma unroll
  for (int g = 0; g < 3; g++) { o.I.v[g] = I[g]; o.acc.v[g] = acc[g]; }
  return o;
}
__device__ __forceinline__ void direction_dispatch_faithful(const LayerSeg& P, bool secL, const double (&cur)[3],
                                                            const double (&upR)[3], const double (&kap)[3],
                                                            const double (&kR)[3], double (&I)[3], double (&acc)[3]) {
  const FaithOut o = direction_call_faithful(P.d[0], P.d[1], P.d[2], P.w, P.kind, secL, Tri{{cur[0], cur[1], cur[2]}},
                                             Tri{{upR[0], upR[1], upR[2]}}, Tri{{kap[0], kap[1], kap[2]}},
                                             Tri{{kR[0], kR[1], kR[2]}}, Tri{{acc[0], acc[1], acc[2]}});
#pragma unroll
  for (int g = 0; g < 3; g++) { I[g] = o.I.v[g]; acc[g] = o.acc.v[g]; }
}

__device__ __forceinline__ void prefetch_l1(const double* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

constexpr double kKappaFloor = 1e-100;  // FAST mode: kappa = 0 is evaluated as this (every formula takes its limit)
constexpr double kTwoM200 = 6.223015277861141707e-61;  // 2^-200

// block = 8 warps; a warp covers 31 cells of one row plus, in lane 0, the recomputed last cell of the strip to its
// left (for strip 0 that is the pad column, which behaves as "no neighbour").
constexpr int kMaxBatch = 40;  // tasks per launch: 40 x 728 B of parameters (the limit is 32764 B since CUDA 12.1)
struct BatchParams {
  StepParams t[kMaxBatch];
};
static_assert(sizeof(BatchParams) + 64 <= 32764, "kernel parameter space");

// gridDim.z tasks (zones) per launch, all at the same layer index: one launch per layer keeps the device full
// (thousands of blocks) instead of many small concurrent kernels.
// A layer whose pattern has a very short segment (a corner clip, len < 1e-2 cell; 1.7% of the (direction, layer)
// pairs) is evaluated with the reference's own operation sequence even in FAST mode: there tau is tiny and the
// rounding noise of the reference's (Iin-Iout)/log(Iin/Iout), ~1.1e-16/tau, would otherwise show up as a parity
// difference.  That branch is warp-uniform (P.thin is a per-layer table entry) and kept out of line.
template <bool FAITHFUL, int EXPV, int MINB>
__global__ void __launch_bounds__(256, MINB)
sweep_cell_kernel(const __grid_constant__ BatchParams bp, int N, int n, int np1, int npl3) {
  // n, np1 = n + 1 and npl3 = 3 (n+1)^2 are the same for every task: as top-level parameters they are constant-bank
  // operands instead of values re-derived from the task table for every direction
  const StepParams& sp = bp.t[blockIdx.z];
  const double* __restrict__ kappa = sp.kappa;
  __shared__ double sT[kExpTableSize];
  // Programmatic dependent launch (set_tuning "pdl"): the next layer's grid may start as soon as every block of this
  // one is running, do its own prologue (table, indices, opacities -- nothing a sweep kernel writes) and then wait at
  // griddepcontrol.wait below until this grid has completed and flushed.  Both instructions are no-ops for a grid
  // launched without the attribute.
  asm volatile("griddepcontrol.launch_dependents;");
  if (threadIdx.y * 32 + threadIdx.x < kExpTableSize) sT[threadIdx.y * 32 + threadIdx.x] = kExpTable32[threadIdx.y * 32 + threadIdx.x];
  __syncthreads();
  const int a = blockIdx.x * 31 - 1 + (int)threadIdx.x, b = blockIdx.y * blockDim.y + threadIdx.y;
  if (b >= n) return;                                          // warp-uniform
  const bool inRow = a < n;                                    // lanes beyond the row only take part in the shuffles
  const bool writer = inRow && threadIdx.x >= 1;
  const bool cell = inRow && a >= 0;                           // a real cell (not the pad column)
  const int laneIsK = sp.laneIsK;
  const int sA = laneIsK ? sp.sk : sp.sj, sB = laneIsK ? sp.sj : sp.sk;
  const int leaf = sp.origin + a * sA + b * sB;
  double kap[3], kapF[3], kR[3];
#pragma unroll
  for (int g = 0; g < 3; g++) {
    const double* kg = kappa + (int64_t)g * N + leaf;
    kapF[g] = cell ? kg[0] : 0.;
    kR[g] = (cell && b > 0) ? kg[-sB] : 0.;                    // kappa = 0 outside: exp(-0) = 1 exactly
    kap[g] = kapF[g] > 0. ? kapF[g] : kKappaFloor;
  }
  const double kmax = fmax(fmax(fmax(kap[0], kap[1]), fmax(kap[2], kR[0])), fmax(kR[1], kR[2]));
  double A[3] = {0., 0., 0.}, acc[3] = {0., 0., 0.};
  const int pidx = (b + 1) * np1 + (inRow ? a + 1 : 0);
  const int ndir = sp.ndir;
  const int dstride = npl3;                                    // one direction's plane (3 groups per cell)
  const int up = 3 * np1;                                      // one row up
  const double* pin = sp.planeIn + 3 * pidx;
  double* pout = sp.planeOut + 3 * pidx;
  asm volatile("griddepcontrol.wait;" ::: "memory");           // the previous layer's planes and accumulators
  for (int q = 0; q < ndir; q++, pin += dstride, pout += dstride) {
    const LayerSeg& P = sp.P[q];
    const int kind = P.kind;
    // Pull the NEXT direction's plane values towards L1 with prefetch instructions: they hold no destination
    // register (at 64 registers per thread a register prefetch is spilled at once, and the spill store then waits
    // for the load -- 35% of all stall samples in the r01 profile).
    if (q + 1 < ndir) {
      prefetch_l1(pin + dstride);
      prefetch_l1(pin + dstride - up);
    }
    double cur[3], upR[3] = {0., 0., 0.}, I[3];
#pragma unroll
    for (int g = 0; g < 3; g++) cur[g] = pin[g];
    // kinds 1,2: second segment fed from k-1, third (kind 2) from j-1; kinds 3,4 the other way round
    const bool secL = (kind <= 2) == (laneIsK != 0);
    if (kind == 2 || kind == 4 || (kind != 0 && !secL)) {
#pragma unroll
      for (int g = 0; g < 3; g++) upR[g] = pin[g - up];
    }
    if (FAITHFUL || P.thin) direction_dispatch_faithful(P, secL, cur, upR, kapF, kR, I, acc);
    else direction_dispatch_fast<EXPV>(P, secL, kmax, cur, upR, kap, kR, I, A, sT);
    if (writer) {
#pragma unroll
      for (int g = 0; g < 3; g++) pout[g] = I[g];
    }
  }
  if (writer) {
#pragma unroll
    for (int g = 0; g < 3; g++) {
      if (!FAITHFUL) acc[g] = fma(A[g], kTwoM200 / kap[g], acc[g]);
      double* p = sp.acc + (int64_t)g * N + leaf;
      *p = sp.firstInSlot ? acc[g] : __dadd_rn(*p, acc[g]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Two cells per thread (FAST arithmetic; set_tuning "cells" = 2, the default): a thread owns the cells (a, b0) and
// (a, b0 + 1) of the layer, b0 even.  What the upper cell needs from the row below it -- the previous layer's plane
// value of (a, b0) and the intensity the lower cell hands over along the row axis -- is already in the thread's
// registers: no recomputed (b-1) segment and no second plane load for the upper cell, and the per-direction
// bookkeeping (table reads, kind dispatch, pointer updates, prefetches) is paid once per two cells.  The arithmetic
// of every cell is unchanged: the handed-over value is bit for bit what the single-cell kernel recomputes.
//   r01 ncu of the single-cell kernel: 318 warp instructions per direction, 194 of them not FP64 (issue-bound).
// ---------------------------------------------------------------------------------------------------------
template <int EXPV, int NSEG, bool SECL, bool GUARD>
__device__ __forceinline__ void direction_fast2(const LayerSeg& P, const double (&cur0)[3], const double (&cur1)[3],
                                                const double (&upR)[3], const double (&kap0)[3], const double (&kap1)[3],
                                                const double (&kR)[3], double (&I0)[3], double (&I1)[3], double (&A0)[3],
                                                double (&A1)[3], const double* __restrict__ T) {
  const unsigned full = 0xffffffffu;
#pragma unroll
  for (int g = 0; g < 3; g++) {
    const double a1 = segment_fast<EXPV, GUARD>(cur0[g], kap0[g] * P.d[0], P.cs[0], T, A0[g]);
    const double b1 = segment_fast<EXPV, GUARD>(cur1[g], kap1[g] * P.d[0], P.cs[0], T, A1[g]);
    I0[g] = a1; I1[g] = b1;
    if (NSEG >= 2) {
      double rup = 0.;  // xy-segment output of the (b0 - 1) cell, recomputed (lower cell only)
      if (!SECL || NSEG == 3) rup = attenuate_fast<EXPV, GUARD>(upR[g], kR[g] * P.d[0], T);
      const double a_in2 = SECL ? __shfl_up_sync(full, a1, 1) : rup;
      const double b_in2 = SECL ? __shfl_up_sync(full, b1, 1) : a1;   // the lower cell's own xy output
      const double a2 = segment_fast<EXPV, GUARD>(a_in2, kap0[g] * P.d[1], P.cs[1], T, A0[g]);
      const double b2 = segment_fast<EXPV, GUARD>(b_in2, kap1[g] * P.d[1], P.cs[1], T, A1[g]);
      I0[g] = a2; I1[g] = b2;
      if (NSEG == 3) {
        double a_in3, b_in3;
        if (SECL) {
          const double x = __shfl_up_sync(full, rup, 1);
          a_in3 = attenuate_fast<EXPV, GUARD>(x, kR[g] * P.d[1], T);
          b_in3 = a2;                                                 // the lower cell's second segment
        } else {
          a_in3 = __shfl_up_sync(full, a2, 1);
          b_in3 = __shfl_up_sync(full, b2, 1);
        }
        I0[g] = segment_fast<EXPV, GUARD>(a_in3, kap0[g] * P.d[2], P.cs[2], T, A0[g]);
        I1[g] = segment_fast<EXPV, GUARD>(b_in3, kap1[g] * P.d[2], P.cs[2], T, A1[g]);
      }
    }
  }
}

template <int EXPV, bool GUARD>
__device__ __forceinline__ void direction_kinds_fast2(const LayerSeg& P, bool secL, const double (&cur0)[3],
                                                      const double (&cur1)[3], const double (&upR)[3],
                                                      const double (&kap0)[3], const double (&kap1)[3],
                                                      const double (&kR)[3], double (&I0)[3], double (&I1)[3],
                                                      double (&A0)[3], double (&A1)[3], const double* __restrict__ T) {
  const int kind = P.kind;
  if (kind == 0) direction_fast2<EXPV, 1, true, GUARD>(P, cur0, cur1, upR, kap0, kap1, kR, I0, I1, A0, A1, T);
  else if (kind == 1 || kind == 3) {
    if (secL) direction_fast2<EXPV, 2, true, GUARD>(P, cur0, cur1, upR, kap0, kap1, kR, I0, I1, A0, A1, T);
    else direction_fast2<EXPV, 2, false, GUARD>(P, cur0, cur1, upR, kap0, kap1, kR, I0, I1, A0, A1, T);
  } else {
    if (secL) direction_fast2<EXPV, 3, true, GUARD>(P, cur0, cur1, upR, kap0, kap1, kR, I0, I1, A0, A1, T);
    else direction_fast2<EXPV, 3, false, GUARD>(P, cur0, cur1, upR, kap0, kap1, kR, I0, I1, A0, A1, T);
  }
}

// block = 8 warps; a warp covers 31 cells (+ the recomputed halo cell in lane 0) of TWO rows
template <int EXPV, int MINB>
__global__ void __launch_bounds__(256, MINB)
sweep_cell2_kernel(const __grid_constant__ BatchParams bp, int N, int n, int np1, int npl3) {
  const StepParams& sp = bp.t[blockIdx.z];
  const double* __restrict__ kappa = sp.kappa;
  __shared__ double sT[kExpTableSize];
  asm volatile("griddepcontrol.launch_dependents;");
  if (threadIdx.y * 32 + threadIdx.x < kExpTableSize) sT[threadIdx.y * 32 + threadIdx.x] = kExpTable32[threadIdx.y * 32 + threadIdx.x];
  __syncthreads();
  const int a = blockIdx.x * 31 - 1 + (int)threadIdx.x, b0 = 2 * (blockIdx.y * blockDim.y + threadIdx.y);
  if (b0 >= n) return;                                         // warp-uniform
  const bool row1 = b0 + 1 < n;                                // warp-uniform: the upper row exists
  const bool inRow = a < n;
  const bool writer0 = inRow && threadIdx.x >= 1, writer1 = writer0 && row1;
  const bool cell0 = inRow && a >= 0, cell1 = cell0 && row1;
  const int laneIsK = sp.laneIsK;
  const int sA = laneIsK ? sp.sk : sp.sj, sB = laneIsK ? sp.sj : sp.sk;
  const int leaf0 = sp.origin + a * sA + b0 * sB;
  double kap0[3], kap1[3], kF0[3], kF1[3], kR[3];
#pragma unroll
  for (int g = 0; g < 3; g++) {
    const double* kg = kappa + (int64_t)g * N + leaf0;
    kF0[g] = cell0 ? kg[0] : 0.;
    kF1[g] = cell1 ? kg[sB] : 0.;
    kR[g] = (cell0 && b0 > 0) ? kg[-sB] : 0.;                  // kappa = 0 outside: exp(-0) = 1 exactly
    kap0[g] = kF0[g] > 0. ? kF0[g] : kKappaFloor;
    kap1[g] = kF1[g] > 0. ? kF1[g] : kKappaFloor;
  }
  const double kmax = fmax(fmax(fmax(fmax(kap0[0], kap0[1]), fmax(kap0[2], kR[0])), fmax(kR[1], kR[2])),
                           fmax(fmax(kap1[0], kap1[1]), kap1[2]));
  double A0[3] = {0., 0., 0.}, A1[3] = {0., 0., 0.}, acc0[3] = {0., 0., 0.}, acc1[3] = {0., 0., 0.};
  const int pidx = (b0 + 1) * np1 + (inRow ? a + 1 : 0);
  const int ndir = sp.ndir;
  const int dstride = npl3;
  const int up = 3 * np1;                                      // one row up in the plane
  const int up1 = row1 ? up : 0;                               // the upper row's plane values (stay in bounds without it)
  const double* pin = sp.planeIn + 3 * pidx;
  double* pout = sp.planeOut + 3 * pidx;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  for (int q = 0; q < ndir; q++, pin += dstride, pout += dstride) {
    const LayerSeg& P = sp.P[q];
    const int kind = P.kind;
    if (q + 1 < ndir) {
      prefetch_l1(pin + dstride);
      prefetch_l1(pin + dstride + up1);
      prefetch_l1(pin + dstride - up);
    }
    double cur0[3], cur1[3], upR[3] = {0., 0., 0.}, I0[3], I1[3];
#pragma unroll
    for (int g = 0; g < 3; g++) { cur0[g] = pin[g]; cur1[g] = pin[g + up1]; }
    const bool secL = (kind <= 2) == (laneIsK != 0);
    if (kind == 2 || kind == 4 || (kind != 0 && !secL)) {
#pragma unroll
      for (int g = 0; g < 3; g++) upR[g] = pin[g - up];
    }
    if (P.thin) {
      // rare (0.2% of the direction-layers): the reference's operation sequence, cell by cell; the upper cell takes
      // the lower one as its (b-1) cell exactly as the single-cell kernel does
      direction_dispatch_faithful(P, secL, cur0, upR, kF0, kR, I0, acc0);
      direction_dispatch_faithful(P, secL, cur1, cur0, kF1, kF0, I1, acc1);
    } else if (__any_sync(0xffffffffu, kmax * P.dmax > 64.)) {
      direction_kinds_fast2<EXPV, true>(P, secL, cur0, cur1, upR, kap0, kap1, kR, I0, I1, A0, A1, sT);
    } else {
      direction_kinds_fast2<EXPV, false>(P, secL, cur0, cur1, upR, kap0, kap1, kR, I0, I1, A0, A1, sT);
    }
    if (writer0) {
#pragma unroll
      for (int g = 0; g < 3; g++) pout[g] = I0[g];
    }
    if (writer1) {
#pragma unroll
      for (int g = 0; g < 3; g++) pout[g + up] = I1[g];
    }
  }
  if (writer0) {
#pragma unroll
    for (int g = 0; g < 3; g++) {
      const double v = fma(A0[g], kTwoM200 / kap0[g], acc0[g]);
      double* p = sp.acc + (int64_t)g * N + leaf0;
      *p = sp.firstInSlot ? v : __dadd_rn(*p, v);
    }
  }
  if (writer1) {
#pragma unroll
    for (int g = 0; g < 3; g++) {
      const double v = fma(A1[g], kTwoM200 / kap1[g], acc1[g]);
      double* p = sp.acc + (int64_t)g * N + leaf0 + sB;
      *p = sp.firstInSlot ? v : __dadd_rn(*p, v);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Persistent layer loop (FAST arithmetic; set_tuning "persistent": 1 on, 0 off, -1 = for small direction shards).
//
// A multi-GPU rank sweeps 3 zones: a layer launch of 3 zone tasks is ~16 us of fp64 work, but costs ~28 us -- its
// 98K threads are 1.3 "waves" of the 75K the device holds, and the hand-over from one launch to the next (drain of
// the whole grid, ramp-up of the next) is paid 256 times per sweep.  Here ONE launch runs the whole sweep: blocks
// take work items (layer, task, tile) in that order from a counter; a tile of layer L waits until the tiles it
// exchanges plane values with -- itself and its 8 neighbours in the task -- have completed layer L-1 (one progress
// word per tile, release/acquire), so layers overlap wherever the dependencies allow and the device stays full.  Items are handed out in dependency order, so the
// oldest unfinished item never waits: no co-residency requirement, no grid-wide barrier.  The planes stay in global
// memory (L2-resident for a shard's ~24 directions); a plane buffer is rewritten every second layer, and the
// acquire load that ends the wait is what makes the SM's L1 forget the older contents.  The per-cell code and the
// order of every sum are those of sweep_cell2_kernel: results are bit-identical to the per-layer launches.
// ---------------------------------------------------------------------------------------------------------
struct PersistTask {
  const double* kappa;   // [3][N] in the task's layout
  double* acc;           // slot accumulator [3][N]
  int64_t planeOff;      // offset of the task's first direction in a plane buffer (doubles)
  int32_t origin, si, sj, sk, ndir, laneIsK, firstInSlot, tileBase;   // tileBase: first work item of the task in a layer
};
struct PersistParams {
  PersistTask t[kMaxBatch];
  const LayerSeg* seg;   // [task][layer][kMaxDirPerTask]
  double *planeA, *planeB;
  int32_t* counters;     // [0] work counter, [1 + task * tilesPerTask + tile] layers the tile has completed; zero at launch
  int32_t ntask, n, np1, npl3, N, gx, gyT, itemsPerLayer, tilesPerTask;
};

template <int EXPV>
__device__ __forceinline__ void persistent_tile(const PersistParams& pp, const PersistTask& T, const LayerSeg* __restrict__ sP,
                                                int layer, int bx, int byWarp, const double* __restrict__ sT) {
  const int n = pp.n, np1 = pp.np1, N = pp.N;
  const int a = bx * 31 - 1 + (int)threadIdx.x, b0 = 2 * byWarp;
  if (b0 >= n) return;                                         // warp-uniform
  const bool row1 = b0 + 1 < n;
  const bool inRow = a < n;
  const bool writer0 = inRow && threadIdx.x >= 1, writer1 = writer0 && row1;
  const bool cell0 = inRow && a >= 0, cell1 = cell0 && row1;
  const int laneIsK = T.laneIsK;
  const int sA = laneIsK ? T.sk : T.sj, sB = laneIsK ? T.sj : T.sk;
  const int leaf0 = T.origin + layer * T.si + a * sA + b0 * sB;
  double kap0[3], kap1[3], kF0[3], kF1[3], kR[3];
#pragma unroll
  for (int g = 0; g < 3; g++) {
    const double* kg = T.kappa + (int64_t)g * N + leaf0;
    kF0[g] = cell0 ? __ldg(kg) : 0.;
    kF1[g] = cell1 ? __ldg(kg + sB) : 0.;
    kR[g] = (cell0 && b0 > 0) ? __ldg(kg - sB) : 0.;
    kap0[g] = kF0[g] > 0. ? kF0[g] : kKappaFloor;
    kap1[g] = kF1[g] > 0. ? kF1[g] : kKappaFloor;
  }
  const double kmax = fmax(fmax(fmax(fmax(kap0[0], kap0[1]), fmax(kap0[2], kR[0])), fmax(kR[1], kR[2])),
                           fmax(fmax(kap1[0], kap1[1]), kap1[2]));
  double A0[3] = {0., 0., 0.}, A1[3] = {0., 0., 0.}, acc0[3] = {0., 0., 0.}, acc1[3] = {0., 0., 0.};
  const int pidx = (b0 + 1) * np1 + (inRow ? a + 1 : 0);
  const int ndir = T.ndir;
  const int dstride = pp.npl3;
  const int up = 3 * np1;
  const int up1 = row1 ? up : 0;
  const double* pin = ((layer & 1) ? pp.planeA : pp.planeB) + T.planeOff + 3 * pidx;
  double* pout = ((layer & 1) ? pp.planeB : pp.planeA) + T.planeOff + 3 * pidx;
  for (int q = 0; q < ndir; q++, pin += dstride, pout += dstride) {
    const LayerSeg& P = sP[q];
    const int kind = P.kind;
    if (q + 1 < ndir) {
      prefetch_l1(pin + dstride);
      prefetch_l1(pin + dstride + up1);
      prefetch_l1(pin + dstride - up);
    }
    double cur0[3], cur1[3], upR[3] = {0., 0., 0.}, I0[3], I1[3];
#pragma unroll
    for (int g = 0; g < 3; g++) { cur0[g] = pin[g]; cur1[g] = pin[g + up1]; }
    const bool secL = (kind <= 2) == (laneIsK != 0);
    if (kind == 2 || kind == 4 || (kind != 0 && !secL)) {
#pragma unroll
      for (int g = 0; g < 3; g++) upR[g] = pin[g - up];
    }
    if (P.thin) {
      direction_dispatch_faithful(P, secL, cur0, upR, kF0, kR, I0, acc0);
      direction_dispatch_faithful(P, secL, cur1, cur0, kF1, kF0, I1, acc1);
    } else if (__any_sync(0xffffffffu, kmax * P.dmax > 64.)) {
      direction_kinds_fast2<EXPV, true>(P, secL, cur0, cur1, upR, kap0, kap1, kR, I0, I1, A0, A1, sT);
    } else {
      direction_kinds_fast2<EXPV, false>(P, secL, cur0, cur1, upR, kap0, kap1, kR, I0, I1, A0, A1, sT);
    }
    if (writer0) {
#pragma unroll
      for (int g = 0; g < 3; g++) pout[g] = I0[g];
    }
    if (writer1) {
#pragma unroll
      for (int g = 0; g < 3; g++) pout[g + up] = I1[g];
    }
  }
  if (writer0) {
#pragma unroll
    for (int g = 0; g < 3; g++) {
      const double v = fma(A0[g], kTwoM200 / kap0[g], acc0[g]);
      double* p = T.acc + (int64_t)g * N + leaf0;
      *p = T.firstInSlot ? v : __dadd_rn(*p, v);
    }
  }
  if (writer1) {
#pragma unroll
    for (int g = 0; g < 3; g++) {
      const double v = fma(A1[g], kTwoM200 / kap1[g], acc1[g]);
      double* p = T.acc + (int64_t)g * N + leaf0 + sB;
      *p = T.firstInSlot ? v : __dadd_rn(*p, v);
    }
  }
}

__device__ __forceinline__ int ld_relaxed_gpu(const int32_t* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int MINB>
__global__ void __launch_bounds__(256, MINB) sweep_persistent_kernel(const __grid_constant__ PersistParams pp, int32_t* err) {
  extern __shared__ double smem[];
  double* sT = smem;                                           // exp table (kExpTableSize entries)
  LayerSeg* sSeg = reinterpret_cast<LayerSeg*>(smem + kExpTableSize);     // [kMaxDirPerTask] tables of the item's task and layer
  __shared__ int sItem;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if (tid < kExpTableSize) sT[tid] = kExpTable32[tid];
  constexpr int segDoubles = (int)(sizeof(LayerSeg) / sizeof(double)) * kMaxDirPerTask;
  const int warps = blockDim.y;
  const int total = pp.n * pp.itemsPerLayer;
  int32_t* prog = pp.counters + 1;
  for (;;) {
    if (tid == 0) sItem = atomicAdd(pp.counters, 1);
    __syncthreads();
    const int item = sItem;
    if (item >= total) break;
    const int layer = item / pp.itemsPerLayer, inLayer = item - layer * pp.itemsPerLayer;
    int task = 0;
    while (task + 1 < pp.ntask && inLayer >= pp.t[task + 1].tileBase) task++;
    const int local = inLayer - pp.t[task].tileBase;
    const int bx = local % pp.gx, by = local / pp.gx;
    for (int i = tid; i < segDoubles; i += blockDim.x * blockDim.y)
      reinterpret_cast<double*>(sSeg)[i] =
          __ldg(reinterpret_cast<const double*>(pp.seg + ((size_t)task * pp.n + layer) * kMaxDirPerTask) + i);
    if (threadIdx.y == 0 && layer > 0) {
      // the tile's inputs are the previous layer's planes of itself and its upstream neighbours, and the buffer it
      // writes was read in the previous layer by itself and its downstream neighbours: wait until those (up to) 9
      // tiles have completed `layer` layers.  Polls are relaxed loads (they leave the SM's L1 alone); the acquire
      // fence after them is what invalidates it.
      if (threadIdx.x < 9) {
        const int nx = bx + (int)(threadIdx.x % 3) - 1, ny = by + (int)(threadIdx.x / 3) - 1;
        if (nx >= 0 && nx < pp.gx && ny >= 0 && ny < pp.gyT) {
          const int32_t* flag = prog + task * pp.tilesPerTask + ny * pp.gx + nx;
          unsigned spins = 0;
          while (ld_relaxed_gpu(flag) < layer) {
            __nanosleep(32);
            if (++spins > (1u << 26)) { atomicExch(err, RTB200_ERR_CUDA); break; }
          }
        }
      }
      __syncwarp();
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
    persistent_tile<1>(pp, pp.t[task], sSeg, layer, bx, by * warps + (int)threadIdx.y, sT);
    __syncthreads();                       // every thread's plane and accumulator stores are issued
    if (tid == 0)
      asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(prog + task * pp.tilesPerTask + local), "r"(layer + 1) : "memory");
  }
}

// z-major layout.  A zone whose sweep axis is the contiguous (z) axis of the leaf order has its lanes run along y:
// every lane of a kappa load or accumulator store would touch its own 32-byte sector (measured: those 8 zones take 1.7x
// the time of the others).  Such tasks read a z-major copy of kappa, index (z*n + x)*n + y, and write their accumulator
// in the same layout, so that lanes are contiguous again; the final merge transposes those slots back through a
// shared-memory tile.
__global__ void transpose_kappa_kernel(const double* __restrict__ kap, double* __restrict__ kapT, int n) {
  __shared__ double tile[32][33];
  const int x = blockIdx.z % n, g = blockIdx.z / n;
  const int64_t N = (int64_t)n * n * n;
  const int z0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {   // read leaf order: z contiguous
    const int y = y0 + r, z = z0 + threadIdx.x;
    if (y < n && z < n) tile[r][threadIdx.x] = kap[g * N + ((int64_t)x * n + y) * n + z];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {   // write z-major: y contiguous
    const int z = z0 + r, y = y0 + threadIdx.x;
    if (y < n && z < n) kapT[g * N + ((int64_t)z * n + x) * n + y] = tile[threadIdx.x][r];
  }
}

// J += sum of the z-major slots [first, first + count), transposed back to leaf order
__global__ void merge_transposed_kernel(const double* __restrict__ acc, int first, int count, int n,
                                        double* __restrict__ J) {
  __shared__ double tile[32][33];
  const int x = blockIdx.z % n, g = blockIdx.z / n;
  const int64_t N = (int64_t)n * n * n, total = 3 * N;
  const int z0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int z = z0 + r, y = y0 + threadIdx.x;
    if (y < n && z < n) {
      const int64_t i = g * N + ((int64_t)z * n + x) * n + y;
      double s = acc[(int64_t)first * total + i];
      for (int k = 1; k < count; k++) s = __dadd_rn(s, acc[(int64_t)(first + k) * total + i]);
      tile[r][threadIdx.x] = s;
    }
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int y = y0 + r, z = z0 + threadIdx.x;
    if (y < n && z < n) {
      double* p = J + g * N + ((int64_t)x * n + y) * n + z;
      *p = __dadd_rn(*p, tile[threadIdx.x][r]);
    }
  }
}

// sum of the slot accumulators -> J (fixed order: slot 0, 1, ...)
__global__ void merge_slots_kernel(const double* __restrict__ acc, int nslots, int64_t total, double* __restrict__ J) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    double s = acc[i];
    for (int k = 1; k < nslots; k++) s = __dadd_rn(s, acc[(int64_t)k * total + i]);
    J[i] = s;
  }
}

__global__ void diffuse_rates_kernel(const double* __restrict__ J, int64_t n, double fourPi, double a0, double a1,
                                     double a2, double b3, double c2, double c3, double* k24, double* k25, double* k26) {
  // equiSources.f90:3546-3553
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double t1 = fourPi * J[i], t2 = fourPi * J[n + i], t3 = fourPi * J[2 * n + i];
    k24[i] = k24[i] + t1 * a0 + t2 * a1 + t3 * a2;
    k25[i] = k25[i] + t3 * b3;
    k26[i] = k26[i] + t2 * c2 + t3 * c3;
  }
}

int launch_diffuse_rates(Context& c, const double* J, const double* ksi24, const double* ksi25, const double* ksi26,
                         double* k24, double* k25, double* k26, cudaStream_t s) {
  int blocks = std::min<int64_t>((c.nleaf + 255) / 256, (int64_t)c.smCount * 16);
  diffuse_rates_kernel<<<blocks, 256, 0, s>>>(J, c.nleaf, 4. * kPi, ksi24[0], ksi24[1], ksi24[2], ksi25[0], ksi26[0],
                                              ksi26[1], k24, k25, k26);
  RTB_CUDA(cudaGetLastError());
  return RTB200_OK;
}

// ---------------------------------------------------------------------------------------------------------
// host side: plan + launch
// ---------------------------------------------------------------------------------------------------------
// FAST mode evaluates a layer with the reference's own operation sequence when one of its segments is shorter than
// this (in cells).  Every segment of a cell enters Jmean with weight 1/(nseg * ndirections) whatever its length
// (transportRoutinesModule.f90:953-955), and the reference formula's rounding noise is ~1.1e-16 / tau_segment, so a
// segment disturbs a cell's Jmean by ~1.1e-16 / (576 tau_segment) where all directions contribute alike, and by up to
// ~1.1e-16 / (3 tau_segment) where a single direction dominates (cells shadowed from most sides).  Measured at 128^3
// (tau_cell >= 1e-3): threshold 1e-4 leaves a worst cell at 1.03e-9 between FAST and FAITHFUL, 1e-3 at ~1e-10.
// 0.2% of the (direction, layer) pairs take this branch.
constexpr double kThinLen = 1e-3;

static LayerSeg make_layer_seg(const RayPattern& p, double cellSize, double weight) {
  LayerSeg L;
  std::memset(&L, 0, sizeof(L));
  double len[3] = {p.xy_len, 0., 0.};
  int kind = 0;
  if (p.xyTop == 1) kind = 0;
  else if (p.yzTop == 1) {  // xy leaves through x = 1 -> yz ray next door
    kind = p.xzActive ? 2 : 1;
    len[1] = p.yz_len;
    len[2] = p.xzActive ? p.xz_len : 0.;
  } else {                  // xy leaves through y = 1 -> xz ray next door
    kind = p.yzActive ? 4 : 3;
    len[1] = p.xz_len;
    len[2] = p.yzActive ? p.yz_len : 0.;
  }
  L.kind = kind;
  L.nseg = 1 + (kind != 0) + (kind == 2 || kind == 4);
  for (int s = 0; s < 3; s++) {
    L.d[s] = cellSize * len[s];
  }
  L.w = weight;
  const double wn = weight / (double)L.nseg;
  L.dmax = std::max(L.d[0], std::max(L.d[1], L.d[2]));
  for (int s = 0; s < 3; s++) L.cs[s] = L.d[s] > 0. ? 1.6069380442589903e60 * wn / L.d[s] : 0.;  // 2^200 wn / d
  L.thin = 0;
  for (int sg = 0; sg < L.nseg; sg++)
    if (len[sg] < kThinLen) L.thin = 1;
  return L;
}


// `warps` = rows of the layer a block covers with one warp each (8, 4 or 2): small direction shards (multi-GPU ranks) put
// only a few hundred 8-warp blocks on the device per layer, 1.5 "waves" of which cost as much as 2; smaller blocks
// spread the same rows evenly over the SMs
template <class Kernel>
static cudaError_t launch_layer(Kernel kern, dim3 grid, cudaStream_t s, bool pdl, const BatchParams& bp, int N, int n,
                                int warps = 8) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(32, warps);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, bp, N, n, n + 1, 3 * (n + 1) * (n + 1));
}

// `pdl`: programmatic dependent launch on the previous launch of the stream (only for a layer whose predecessor in
// the stream is the previous layer's sweep kernel)
static cudaError_t launch_cells(int dense, int expv, bool faithful, bool pdl, dim3 grid, cudaStream_t s,
                                const BatchParams& bp, int N, int n, int cells, int warps, int smCount) {
  // `dense`: 0 = compiler's choice of registers (2 blocks per SM), 1 = cap for 3 blocks, 2 = cap for 4 blocks
  // automatic: two cells per thread pay (1-2%) when the launch has many zone tasks; a small direction shard (3 zone
  // tasks on each of 8 GPUs) is 1.3 waves of two-cell threads but fills the device evenly with one-cell threads
  // (measured on the shards of an 8-GPU run: 7.03 against 7.21 ms mean, 7.31 against 7.93 ms for the slowest rank)
  if (cells == 0) cells = (n >= 192 && grid.z >= 6) ? 2 : 1;
  if (faithful || expv != 1) cells = 1;
  const int rowsTotal = cells == 2 ? (n + 1) / 2 : n;           // rows of threads per task and layer
  if (warps != 8 && warps != 4 && warps != 2) {
    // automatic: the largest block that still gives the device ~6 blocks per resident 8-warp slot
    const long long slots = (long long)smCount * (cells == 2 ? 2 : 4);
    warps = 8;
    while (warps > 2 && (long long)grid.x * ((rowsTotal + warps - 1) / warps) * grid.z < 6 * slots) warps /= 2;
  }
  grid.y = (unsigned)((rowsTotal + warps - 1) / warps);
  if (faithful) return launch_layer(sweep_cell_kernel<true, 0, 2>, grid, s, pdl, bp, N, n, warps);
  if (cells == 2) {   // two cells per thread
    // register budget: default 2 blocks of 256 threads per SM (119 registers, no spills; as many cells in flight as
    // the single-cell kernel at 4 blocks); "dense" 3 / 4 cap the registers for 3 / 4 blocks
    if (dense >= 4) return launch_layer(sweep_cell2_kernel<1, 4>, grid, s, pdl, bp, N, n, warps);
    if (dense == 3) return launch_layer(sweep_cell2_kernel<1, 3>, grid, s, pdl, bp, N, n, warps);
    return launch_layer(sweep_cell2_kernel<1, 2>, grid, s, pdl, bp, N, n, warps);
  }
#define RTB_LAUNCH(E)                                                                                    \
  if (dense == 1) return launch_layer(sweep_cell_kernel<false, E, 3>, grid, s, pdl, bp, N, n, warps);    \
  else if (dense >= 2) return launch_layer(sweep_cell_kernel<false, E, 4>, grid, s, pdl, bp, N, n, warps); \
  else return launch_layer(sweep_cell_kernel<false, E, 2>, grid, s, pdl, bp, N, n, warps)
  if (expv == 1) { RTB_LAUNCH(1); } else { RTB_LAUNCH(0); }
#undef RTB_LAUNCH
}


// Runs the sweep as one launch (see sweep_persistent_kernel).  Returns -1 when it does not apply.
static int run_persistent(Context& c, int n, const double* uvb, double* dJout, cudaStream_t s, int ndir) {
  const int64_t N = c.nleaf, npl = (int64_t)(n + 1) * (n + 1);
  const int ntask = (int)c.uniTasks.size();
  if (ntask < 1 || ntask > kMaxBatch || c.uniSlots < ntask) return -1;   // every task needs its own accumulator slot
  const int warps = (c.tune.blockWarps == 8 || c.tune.blockWarps == 4) ? c.tune.blockWarps : 2;   // small tiles by default
  const int gx = (n + 30) / 31, rowsTotal = (n + 1) / 2, gyT = (rowsTotal + warps - 1) / warps;
  static thread_local PersistParams pp;
  // per-layer tables of all tasks on the device (kept while the plan is the same)
  const size_t segPerTask = (size_t)n * kMaxDirPerTask;
  if (c.marchSegKey != c.uniPlanKey) {
    if (int e = ensure_buffer((void**)&c.dMarchSeg, &c.marchSegBytes, (size_t)ntask * segPerTask * sizeof(LayerSeg))) return e;
    for (int t = 0; t < ntask; t++)
      RTB_CUDA(cudaMemcpyAsync((LayerSeg*)c.dMarchSeg + (size_t)t * segPerTask, c.uniTasks[t].seg.data(),
                               segPerTask * sizeof(LayerSeg), cudaMemcpyHostToDevice, s));
    RTB_CUDA(cudaStreamSynchronize(s));
    c.marchSegKey = c.uniPlanKey;
  }
  const size_t counterBytes = ((size_t)ntask * gx * gyT + 1) * sizeof(int32_t);
  if (int e = ensure_buffer((void**)&c.dMarchProg, &c.marchProgBytes, counterBytes)) return e;
  const size_t smemBytes = kExpTableSize * sizeof(double) + (size_t)kMaxDirPerTask * sizeof(LayerSeg);
  auto kern = sweep_persistent_kernel<2>;
  if (smemBytes > 48 * 1024) RTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
  int perSm = 0;
  RTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kern, 32 * warps, smemBytes));
  if (perSm < 1) return -1;
  int items = 0;
  for (int t = 0; t < ntask; t++) {
    const UniTaskHost& T = c.uniTasks[t];
    PersistTask& q = pp.t[t];
    q.kappa = T.transposed ? c.dKappaT : c.dKappa;
    q.acc = c.dAcc + (size_t)T.slot * 3 * N;
    q.planeOff = (int64_t)T.planeFirst * 3 * npl;
    q.origin = (int32_t)T.origin; q.si = (int32_t)T.si; q.sj = (int32_t)T.sj; q.sk = (int32_t)T.sk;
    q.ndir = T.ndir; q.laneIsK = T.laneIsK; q.firstInSlot = T.firstInSlot; q.tileBase = items;
    items += gx * gyT;
  }
  pp.seg = (const LayerSeg*)c.dMarchSeg;
  pp.planeA = c.dPlanes; pp.planeB = c.dPlanes + (size_t)ndir * 3 * npl;
  pp.counters = c.dMarchProg;
  pp.ntask = ntask; pp.n = n; pp.np1 = n + 1; pp.npl3 = (int32_t)(3 * npl); pp.N = (int32_t)N; pp.gx = gx; pp.gyT = gyT;
  pp.itemsPerLayer = items;
  pp.tilesPerTask = gx * gyT;
  const int blocks = (int)std::min<int64_t>((int64_t)perSm * c.smCount, (int64_t)items * n);
  RTB_CUDA(cudaEventRecord(c.evSweep0, s));
  if (c.uniStdSlots < ntask && c.uniStdSlots < c.uniSlots) {
    dim3 tg((n + 31) / 32, (n + 31) / 32, 3 * n);
    transpose_kappa_kernel<<<tg, dim3(32, 8), 0, s>>>(c.dKappa, c.dKappaT, n);
  }
  {
    const int64_t total = (int64_t)2 * ndir * 3 * npl;
    int fb = (int)std::min<int64_t>((total + 255) / 256, (int64_t)c.smCount * 16);
    fill_planes_kernel<<<fb, 256, 0, s>>>(c.dPlanes, npl, total, uvb[0], uvb[1], uvb[2]);
  }
  RTB_CUDA(cudaMemsetAsync(c.dMarchProg, 0, counterBytes, s));
  kern<<<dim3(blocks), dim3(32, warps), smemBytes, s>>>(pp, c.dErr);
  RTB_CUDA(cudaEventRecord(c.evSweep1, s));
  {
    const int nStd = std::min(c.uniStdSlots, c.uniSlots);
    int mb = std::min<int64_t>((3 * N + 255) / 256, (int64_t)c.smCount * 16);
    if (nStd > 0) merge_slots_kernel<<<mb, 256, 0, s>>>(c.dAcc, nStd, 3 * N, dJout);
    else RTB_CUDA(cudaMemsetAsync(dJout, 0, 3 * N * sizeof(double), s));
    if (nStd < c.uniSlots) {
      dim3 tg((n + 31) / 32, (n + 31) / 32, 3 * n);
      merge_transposed_kernel<<<tg, dim3(32, 8), 0, s>>>(c.dAcc, nStd, c.uniSlots - nStd, n, dJout);
    }
  }
  RTB_CUDA(cudaGetLastError());
  c.uniLaunches = 4;
  c.lastSweepLaunches = 1;
  c.lastLaunches = 7;
  return RTB200_OK;
}

int diffuse_uniform(Context& c, int nAngularLevel, const double* uvb, const std::vector<Direction>& dirs,
                    double* dJout, cudaStream_t s, int64_t* nsegOut) {
  const int n = c.nx;
  const int64_t N = c.nleaf, nn = (int64_t)n * n, npl = (int64_t)(n + 1) * (n + 1);
  const int ndir = (int)dirs.size();
  const int64_t nraysTotal = 12LL << (2 * (nAngularLevel - 1));
  const double weight = (double)(1.f / (float)nraysTotal);  // equiSources.f90:1386 (single-precision division)
  const double cellSize = c.boxSize / (double)n;             // equiSources.f90:1570

  // ---- plan: pattern tables and zone tasks.  Depends only on the grid size and the direction list, so it is
  //      cached across the outer transport<->chemistry iterations. ----
  std::string planKey;
  {
    char buf[128];
    snprintf(buf, sizeof(buf), "%d:%a:%d:%d:%d:%d:%d:%d:%d:", n, c.boxSize, nAngularLevel, c.tune.slots, c.tune.lockstep, c.tune.dirsPerTask,
             c.tune.transposeZ, 0, c.tune.blockWarps);
    planKey = buf;
    for (const auto& d : dirs) { snprintf(buf, sizeof(buf), "%lld,", (long long)d.iray); planKey += buf; }
  }
  if (planKey != c.uniPlanKey) {
    c.uniTasks.clear();
    std::vector<RayPattern> pat;
    int64_t nseg = 0;
    int planeCursor = 0;
    // Directions per task.  Every task contributes (n/31)*(n/8) blocks per layer launch whatever its direction
    // count, so a shard with few zones (multi-GPU runs) fills the device better when its zones are cut into smaller
    // pieces: pick the piece size that minimises (waves of resident blocks) x (block time ~ directions + fixed part).
    int dpt = std::max(1, std::min(c.tune.dirsPerTask > 0 ? c.tune.dirsPerTask : kMaxDirPerTask, kMaxDirPerTask));
    if (c.tune.dirsPerTask <= 0) {
      int perZone[25] = {0};
      for (int d = 0; d < ndir; d++) perZone[dirs[d].izone]++;
      const int64_t bpt = (int64_t)((n + 30) / 31) * ((n + 7) / 8);
      const int64_t cap = (int64_t)c.smCount * (c.tune.minBlocks >= 2 ? 4 : (c.tune.minBlocks == 1 ? 3 : 2));
      double best = 1e300;
      for (int d = kMaxDirPerTask; d >= 2; d--) {
        int64_t T = 0;
        int maxPiece = 0;
        for (int z = 1; z <= 24; z++)
          if (perZone[z]) {
            const int pieces = (perZone[z] + d - 1) / d;
            T += pieces;
            maxPiece = std::max(maxPiece, (perZone[z] + pieces - 1) / pieces);
          }
        if (T == 0 || T > kMaxBatch) continue;
        // with the block size chosen per launch (block_warps = 0) a partly filled last wave costs its share only
        const double fill = (double)(T * bpt) / (double)cap;
        const double waves = c.tune.blockWarps == 0 ? std::max(1.0, fill) : std::ceil(fill);
        const double cost = waves * (maxPiece + 1.5);
        if (cost < best * 0.97) { best = cost; dpt = d; }   // prefer larger pieces unless clearly better
      }
    }
    for (int z = 1; z <= 24; z++) {
      std::vector<int> mine;
      for (int d = 0; d < ndir; d++)
        if (dirs[d].izone == z) mine.push_back(d);
      if (mine.empty()) continue;
      const int pieces = ((int)mine.size() + dpt - 1) / dpt;
      size_t o = 0;
      for (int pc = 0; pc < pieces; pc++) {
        const size_t cnt = (mine.size() - o + (pieces - pc) - 1) / (pieces - pc);  // balanced split
        UniTaskHost T;
        ZoneStrides zs = zone_strides(z, n);
        T.origin = zs.origin; T.si = zs.stride[0]; T.sj = zs.stride[1]; T.sk = zs.stride[2];
        T.izone = z;
        T.ndir = (int)cnt;
        T.planeFirst = planeCursor;
        planeCursor += T.ndir;
        int64_t aj = T.sj < 0 ? -T.sj : T.sj, ak = T.sk < 0 ? -T.sk : T.sk;
        T.laneIsK = ak <= aj;
        T.seg.assign((size_t)n * kMaxDirPerTask, LayerSeg());
        for (int q = 0; q < T.ndir; q++) {
          const Direction& dd = dirs[mine[o + q]];
          if (dd.status) return dd.status;
          layer_patterns_level0(dd.phi, dd.theta, n, pat);
          for (int i = 0; i < n; i++) {
            if (pat[i].status) return pat[i].status;
            LayerSeg L = make_layer_seg(pat[i], cellSize, weight);
            T.seg[(size_t)i * kMaxDirPerTask + q] = L;
            nseg += (int64_t)L.nseg * nn;
          }
        }
        o += cnt;
        c.uniTasks.push_back(std::move(T));
      }
    }
    const int ntask = (int)c.uniTasks.size();
    // slots = zones in flight, each an independent stream / graph branch with a private J accumulator;
    // longest-processing-time-first assignment of the tasks to the slots
    int slots = c.tune.slots;
    if (slots <= 0) slots = c.tune.lockstep ? kMaxBatch : 24;
    slots = std::max(1, std::min(slots, ntask));
    std::vector<int> order(ntask), load(slots, 0), seen(slots, 0);
    for (int t = 0; t < ntask; t++) order[t] = t;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return c.uniTasks[x].ndir > c.uniTasks[y].ndir; });
    if (c.tune.lockstep) {
      slots = std::min(slots, kMaxBatch);
      for (int t = 0; t < ntask; t++) c.uniTasks[t].slot = t % slots;   // batch b = tasks [b*slots, (b+1)*slots)
    } else {
      for (int t : order) {
        int best = 0;
        for (int k = 1; k < slots; k++)
          if (load[k] < load[best]) best = k;
        c.uniTasks[t].slot = best;
        load[best] += c.uniTasks[t].ndir;
      }
    }
    c.uniStdSlots = slots;
    if (c.tune.transposeZ && c.tune.lockstep && ntask <= slots && n >= 32) {
      // every task owns its slot: tasks sweeping along the contiguous axis switch to the z-major layout (see
      // transpose_kappa_kernel) and are moved behind the others so that their slots form one range
      std::stable_partition(c.uniTasks.begin(), c.uniTasks.end(), [](const UniTaskHost& T) { return T.si != 1 && T.si != -1; });
      int nStd = 0;
      for (int t = 0; t < ntask; t++) {
        UniTaskHost& T = c.uniTasks[t];
        T.slot = t;
        if (T.si != 1 && T.si != -1) { nStd++; continue; }
        const int64_t physT[3] = {n, 1, (int64_t)n * n};
        ZoneStrides zs = zone_strides_layout(T.izone, n, physT);
        T.origin = zs.origin; T.si = zs.stride[0]; T.sj = zs.stride[1]; T.sk = zs.stride[2];
        const int64_t aj = T.sj < 0 ? -T.sj : T.sj, ak = T.sk < 0 ? -T.sk : T.sk;
        T.laneIsK = ak <= aj;
        T.transposed = 1;
      }
      c.uniStdSlots = nStd;
      if (nStd < ntask)
        if (int st = ensure_buffer((void**)&c.dKappaT, &c.kappaTBytes, (size_t)3 * N * sizeof(double))) return st;
    }
    for (int t = 0; t < ntask; t++) {  // first task of a slot (in issue order) overwrites the accumulator
      c.uniTasks[t].firstInSlot = !seen[c.uniTasks[t].slot];
      seen[c.uniTasks[t].slot] = 1;
    }
    if (int st = ensure_buffer((void**)&c.dAcc, &c.accBytes, (size_t)slots * 3 * N * sizeof(double))) return st;
    if (int st = ensure_buffer((void**)&c.dPlanes, &c.planeBytes, (size_t)2 * std::max(ndir, 1) * 3 * npl * sizeof(double))) return st;
    c.uniPlanKey = planKey;
    c.uniSlots = slots;
    c.uniNseg = nseg;
    if (c.graphExec) { cudaGraphExecDestroy(c.graphExec); c.graphExec = nullptr; }
  }
  const int ntask = (int)c.uniTasks.size(), slots = c.uniSlots;
  if (nsegOut) *nsegOut = c.uniNseg;
  if (ntask == 0) {
    RTB_CUDA(cudaMemsetAsync(dJout, 0, 3 * N * sizeof(double), s));
    c.lastSweepLaunches = 0;
    c.lastLaunches = 1;
    return RTB200_OK;
  }

  if (int st = ensure_buffer((void**)&c.dAcc, &c.accBytes, (size_t)c.uniSlots * 3 * N * sizeof(double))) return st;
  if (int st = ensure_buffer((void**)&c.dPlanes, &c.planeBytes, (size_t)2 * std::max(ndir, 1) * 3 * npl * sizeof(double))) return st;
  {
    // one launch for the whole sweep: on request only.  Measured on one B200 (profiles/r02h_*): the shards of an
    // 8-GPU run take 6.99-7.22 ms mean / 7.26-7.50 ms max against 7.03 / 7.31 ms with per-layer launches of one-cell
    // threads, and 4-, 2- and 1-GPU shards are 9-16% SLOWER (the layer tables come from shared memory instead of the
    // constant bank, 128 registers with spills, an L1 invalidation per tile), so per-layer launches stay the default
    const bool fast = c.mathMode != RTB200_MATH_FAITHFUL && c.tune.expVariant == 1 && c.tune.lockstep;
    const bool want = c.tune.persistent == 1 || (c.tune.persistent < 0 && ntask <= 6 && n >= 64);
    if (fast && want) {
      const int pst = run_persistent(c, n, uvb, dJout, s, ndir);
      if (pst >= 0) return pst;
    }
  }

  dim3 grid((n + 30) / 31, (n + 7) / 8, 1);
  double* planeA = c.dPlanes;
  double* planeB = c.dPlanes + (size_t)ndir * 3 * npl;
  const bool faithful = c.mathMode == RTB200_MATH_FAITHFUL;
  while ((int)c.chainStreams.size() < slots) {
    cudaStream_t cs;
    RTB_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    cudaEvent_t ce;
    RTB_CUDA(cudaEventCreateWithFlags(&ce, cudaEventDisableTiming));
    c.chainStreams.push_back(cs);
    c.chainEvents.push_back(ce);
  }
  int64_t nLaunched = 1;
  const bool lockstep = c.tune.lockstep != 0;

  // Every task (zone) is a chain of n layer steps; the tasks of one slot run back to back because they share the
  // slot's J accumulator.  Each slot is an independent stream (a parallel branch of the captured graph), so the
  // hardware block scheduler balances the zones and there is no device-wide barrier between layers.
  auto issue = [&](cudaStream_t st) -> int {
    if (c.uniStdSlots < (int)c.uniTasks.size() && c.uniStdSlots < c.uniSlots) {
      dim3 tg((n + 31) / 32, (n + 31) / 32, 3 * n);
      transpose_kappa_kernel<<<tg, dim3(32, 8), 0, st>>>(c.dKappa, c.dKappaT, n);
    }
    {  // both plane buffers start as "boundary intensity everywhere": first-layer input and the permanent pads
      const int64_t total = (int64_t)2 * ndir * 3 * npl;
      int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)c.smCount * 16);
      fill_planes_kernel<<<blocks, 256, 0, st>>>(c.dPlanes, npl, total, uvb[0], uvb[1], uvb[2]);
    }
    // fill one task's parameters for one layer
    auto fill = [&](StepParams& sp, const UniTaskHost& T, int step, int first) -> int {
      sp.acc = c.dAcc + (size_t)T.slot * 3 * N;
      sp.kappa = T.transposed ? c.dKappaT : c.dKappa;
      sp.sj = (int32_t)T.sj; sp.sk = (int32_t)T.sk;
      sp.laneIsK = T.laneIsK; sp.n = n;
      sp.planeIn = ((step & 1) ? planeA : planeB) + (size_t)T.planeFirst * 3 * npl;
      sp.planeOut = ((step & 1) ? planeB : planeA) + (size_t)T.planeFirst * 3 * npl;
      sp.origin = (int32_t)(T.origin + step * T.si);
      for (int q = 0; q < T.ndir; q++) sp.P[q] = T.seg[(size_t)step * kMaxDirPerTask + q];
      sp.ndir = T.ndir;
      sp.firstInSlot = first;
      return T.ndir;
    };
    static thread_local BatchParams bp;
    if (lockstep) {
      // all tasks of a batch advance one layer per launch; each task of a batch has its own accumulator slot
      for (int base = 0; base < ntask; base += slots) {
        const int nb = std::min(slots, ntask - base);
        for (int step = 0; step < n; step++) {
          for (int z = 0; z < nb; z++) {
            const UniTaskHost& T = c.uniTasks[base + z];
            fill(bp.t[z], T, step, T.firstInSlot);  // every layer touches its own cells once per task
          }
          dim3 gz = grid;
          gz.z = nb;
          RTB_CUDA(launch_cells(c.tune.minBlocks, c.tune.expVariant, faithful, c.tune.pdl && step > 0, gz, st, bp, (int)N, n, c.tune.cells, c.tune.blockWarps, c.smCount));
          nLaunched++;
        }
      }
      return RTB200_OK;
    }
    RTB_CUDA(cudaEventRecord(c.evFork, st));
    for (int k = 0; k < slots; k++) {
      cudaStream_t cs = slots > 1 ? c.chainStreams[k] : st;
      if (slots > 1) RTB_CUDA(cudaStreamWaitEvent(cs, c.evFork, 0));
      for (int t = 0; t < ntask; t++) {
        const UniTaskHost& T = c.uniTasks[t];
        if (T.slot != k) continue;
        for (int step = 0; step < n; step++) {
          fill(bp.t[0], T, step, T.firstInSlot);
          RTB_CUDA(launch_cells(c.tune.minBlocks, c.tune.expVariant, faithful, c.tune.pdl && step > 0, grid, cs, bp, (int)N, n, c.tune.cells, c.tune.blockWarps, c.smCount));
          nLaunched++;
        }
      }
      if (slots > 1) {
        RTB_CUDA(cudaEventRecord(c.chainEvents[k], cs));
        RTB_CUDA(cudaStreamWaitEvent(st, c.chainEvents[k], 0));
      }
    }
    return RTB200_OK;
  };
  RTB_CUDA(cudaEventRecord(c.evSweep0, s));
  if (c.tune.useGraph) {
    // The launch sequence depends only on the plan, the mode and the buffers: capture once, replay afterwards.
    char key[256];
    snprintf(key, sizeof(key), "u:%d:%d:%d:%d:%p:%p:%p:%a:%a:%a", n, ntask, slots,
             (int)faithful * 64 + c.tune.minBlocks * 4 + c.tune.expVariant + 1000 * c.tune.lockstep + 10000 * c.tune.pdl +
                 100000 * c.tune.cells + 1000000 * c.tune.blockWarps,
             (void*)c.dAcc, (void*)c.dPlanes, (void*)c.dKappa, uvb[0], uvb[1], uvb[2]);
    if (!c.graphExec || c.graphKey != key) {
      if (c.graphExec) { cudaGraphExecDestroy(c.graphExec); c.graphExec = nullptr; }
      cudaGraph_t graph;
      // capture on the internal stream (the caller's stream may be the legacy default stream, which cannot capture)
      RTB_CUDA(cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeThreadLocal));
      int ist = issue(c.stream);
      cudaError_t ce = cudaStreamEndCapture(c.stream, &graph);
      if (ist) return ist;
      RTB_CUDA(ce);
      RTB_CUDA(cudaGraphInstantiate(&c.graphExec, graph, 0));
      cudaGraphDestroy(graph);
      c.graphKey = key;
    }
    RTB_CUDA(cudaGraphLaunch(c.graphExec, s));
  } else {
    if (int ist = issue(s)) return ist;
  }
  RTB_CUDA(cudaEventRecord(c.evSweep1, s));
  {
    const int nStd = std::min(c.uniStdSlots, slots);
    int blocks = std::min<int64_t>((3 * N + 255) / 256, (int64_t)c.smCount * 16);
    if (nStd > 0) merge_slots_kernel<<<blocks, 256, 0, s>>>(c.dAcc, nStd, 3 * N, dJout);
    else RTB_CUDA(cudaMemsetAsync(dJout, 0, 3 * N * sizeof(double), s));
    if (nStd < slots) {
      dim3 tg((n + 31) / 32, (n + 31) / 32, 3 * n);
      merge_transposed_kernel<<<tg, dim3(32, 8), 0, s>>>(c.dAcc, nStd, slots - nStd, n, dJout);
    }
  }
  RTB_CUDA(cudaGetLastError());
  if (nLaunched > 1) c.uniLaunches = nLaunched;  // a replayed graph issues what was captured
  c.lastSweepLaunches = c.uniLaunches;
  c.lastLaunches = c.uniLaunches + 2;  // + compute_opacities + merge
  return RTB200_OK;
}

}  // namespace rtb
