// Diffuse sweep on a uniform (single-level) grid: replaces the direction loop of equiSources.f90:1389-1806 for
// grids without refinement (configs 2 and 4 of BASELINE.json).
//
// Formulation (DESIGN.md "uniform sweep"):
//  * All directions of one zone share the index rotation (rotateIndicesModule.f90), so they are swept TOGETHER,
//    layer by layer along the zone's sweep axis: one thread owns one cell of the layer, reads kappa1..3 once,
//    loops over the zone's directions and adds all their contributions to J in registers -> one J update per
//    cell per zone instead of one per direction.
//  * Within a layer every cell carries the same 1..3 segments (its "pattern"); a characteristic that leaves a
//    cell sideways continues in the k+1 and/or j+1 neighbour of the SAME layer.  That in-layer hand-over goes
//    through shared memory: phase 1 computes every cell's bottom-entering (xy) segment, phase 2 the segment fed by
//    a neighbour's phase-1 result, phase 3 the one fed by a phase-2 result.  Tiles overlap by one cell on the
//    upstream sides (the halo cells are recomputed) so that no inter-block communication is needed.
//  * The only inter-layer state is the intensity leaving each cell through its top face: one plane of
//    3 doubles per cell per direction, ping-ponged in global memory and meant to stay L2-resident
//    (the number of zones in flight is chosen from the L2 budget).
//  * Zones run in `slots` concurrent lanes (gridDim.z); each slot owns a private J accumulator, so no atomics and a
//    fixed summation order.  A final kernel sums the slot accumulators into J.
#include <algorithm>
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
  int32_t pl[kMaxDirPerTask];  // plane index (within the task) of the q-th direction handled by this launch
  const double* planeIn;   // this task's planes of the previous layer  [ndir][3][n+1][n+1] (padded, see below)
  double* planeOut;
  double* acc;             // slot accumulator [3][N]
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
    int g = (int)((i / perGroup) % 3);
    planes[i] = g == 0 ? u0 : (g == 1 ? u1 : u2);
  }
}

template <bool FAITHFUL, int EXPV>
__device__ __forceinline__ double attenuate(double Iin, double kappa, double dpath, const double* __restrict__ T) {
  if (FAITHFUL) return __dmul_rn(Iin, exp(-__dmul_rn(kappa, dpath)));
  return Iin * exp_neg_only<EXPV>(kappa * dpath, T);
}

// One direction of one cell, straight-line for a given segment count and chain order.
//   NSEG: 1..3 segments.  SECL: the second segment is fed by the neighbour along the LANE axis (cell a-1, same row)
//   and the third by the neighbour along the ROW axis (cell b-1, same lane); !SECL: the other way round.
// Lane-axis hand-over = warp shuffle of the neighbour lane's own result; row-axis hand-over = recompute of what the
// (b-1) cell emits from the previous layer's plane value `upR` and its kappa `kR` (bit-identical to that cell's own
// arithmetic).  Every lane of the warp must call this (shuffles), including the halo lane and out-of-domain lanes.
template <bool FAITHFUL, int EXPV, int NSEG, bool SECL>
__device__ __forceinline__ void direction_body(const LayerSeg& P, const double (&cur)[3], const double (&upR)[3],
                                               const double (&kap)[3], const double (&invk)[3], const double (&kR)[3],
                                               double (&I)[3], double (&acc)[3], const double* __restrict__ T) {
  const unsigned full = 0xffffffffu;
#pragma unroll
  for (int g = 0; g < 3; g++) {
    SegResult r1 = segment_update<FAITHFUL, EXPV>(cur[g], kap[g], P.d[0], invk[g] * P.invd[0], T);
    double Jsum = r1.J;
    I[g] = r1.Iout;
    if (NSEG >= 2) {
      double rup = 0.;  // xy-segment output of the (b-1) cell
      if (!SECL || NSEG == 3) rup = attenuate<FAITHFUL, EXPV>(upR[g], kR[g], P.d[0], T);
      const double in2 = SECL ? __shfl_up_sync(full, r1.Iout, 1) : rup;
      SegResult r2 = segment_update<FAITHFUL, EXPV>(in2, kap[g], P.d[1], invk[g] * P.invd[1], T);
      I[g] = r2.Iout;
      double J3 = 0.;
      if (NSEG == 3) {
        double in3;
        if (SECL) {
          // third segment from the (b-1) cell's SECOND segment, which was fed by the (b-1, a-1) cell's xy segment
          const double x = __shfl_up_sync(full, rup, 1);
          in3 = attenuate<FAITHFUL, EXPV>(x, kR[g], P.d[1], T);
        } else {
          in3 = __shfl_up_sync(full, r2.Iout, 1);  // the (a-1) cell's second segment
        }
        SegResult r3 = segment_update<FAITHFUL, EXPV>(in3, kap[g], P.d[2], invk[g] * P.invd[2], T);
        I[g] = r3.Iout;
        J3 = r3.J;
      }
      Jsum = FAITHFUL ? 0. : (Jsum + r2.J) + J3;
      if (FAITHFUL) {
        // the reference sums the segments in the order xy, xz, yz (transportRoutinesModule.f90:698,818,941);
        // P.kind <= 2 <=> the second segment is the yz ray
        const bool yzSecond = P.kind <= 2;
        const double Jxz = yzSecond ? J3 : r2.J, Jyz = yzSecond ? r2.J : J3;
        Jsum = r1.J;
        if (NSEG == 3 || !yzSecond) Jsum = __dadd_rn(Jsum, Jxz);
        if (NSEG == 3 || yzSecond) Jsum = __dadd_rn(Jsum, Jyz);
      }
    }
    if (FAITHFUL) acc[g] = __dadd_rn(acc[g], __dmul_rn(__ddiv_rn(Jsum, (double)NSEG), P.w));
    else acc[g] = fma(Jsum, P.wn, acc[g]);
  }
}

// block = 8 warps; a warp covers 31 cells of one row plus, in lane 0, the recomputed last cell of the strip to its
// left (for strip 0 that is the pad column, which behaves as "no neighbour").
constexpr int kMaxBatch = 40;  // tasks per launch: 40 x 728 B of parameters (the limit is 32764 B since CUDA 12.1)
struct BatchParams {
  StepParams t[kMaxBatch];
};

// gridDim.z tasks (zones) per launch, all at the same layer index: one launch per layer keeps the device full
// (thousands of blocks) instead of many small concurrent kernels.
template <bool FAITHFUL, int EXPV, int MINB>
__global__ void __launch_bounds__(256, MINB)
sweep_cell_kernel(const __grid_constant__ BatchParams bp, const double* __restrict__ kappa, int N) {
  const StepParams& sp = bp.t[blockIdx.z];
  __shared__ double sT[16];
  if (threadIdx.y == 0 && threadIdx.x < 16) sT[threadIdx.x] = kExpTable[threadIdx.x];
  __syncthreads();
  const int n = sp.n;
  const int a = blockIdx.x * 31 - 1 + (int)threadIdx.x, b = blockIdx.y * 8 + threadIdx.y;
  if (b >= n) return;                                          // warp-uniform
  const bool inRow = a < n;                                    // lanes beyond the row only take part in the shuffles
  const bool writer = inRow && threadIdx.x >= 1;
  const bool cell = inRow && a >= 0;                           // a real cell (not the pad column)
  const int laneIsK = sp.laneIsK;
  const int sA = laneIsK ? sp.sk : sp.sj, sB = laneIsK ? sp.sj : sp.sk;
  const int leaf = sp.origin + a * sA + b * sB;
  const int np1 = n + 1;
  const int npl = np1 * np1;                                   // doubles per padded plane
  double kap[3], invk[3], kR[3];
#pragma unroll
  for (int g = 0; g < 3; g++) {
    const double* kg = kappa + (int64_t)g * N + leaf;
    kap[g] = cell ? kg[0] : 0.;
    kR[g] = (cell && b > 0) ? kg[-sB] : 0.;                    // kappa = 0 outside: exp(-0) = 1 exactly
    if (!FAITHFUL) {
      kap[g] = kap[g] > 0. ? kap[g] : 1e-200;  // kappa = 0 limit through the same formulas (segment_math.cuh)
      invk[g] = 1.0 / kap[g];
    } else {
      invk[g] = 0.;
    }
  }
  double acc[3] = {0., 0., 0.};
  const int pidx = (b + 1) * np1 + (inRow ? a + 1 : 0);
  const int ndir = sp.ndir;
  double cur[3];
  {
    const double* p0 = sp.planeIn + (int64_t)sp.pl[0] * (3 * npl) + pidx;
#pragma unroll
    for (int g = 0; g < 3; g++) cur[g] = p0[g * npl];
  }
  for (int q = 0; q < ndir; q++) {
    const LayerSeg& P = sp.P[q];
    const double* pin = sp.planeIn + (int64_t)sp.pl[q] * (3 * npl) + pidx;
    double* pout = sp.planeOut + (int64_t)sp.pl[q] * (3 * npl) + pidx;
    const int kind = P.kind;
    // issue every load of this direction, and the next direction's own plane values, before the arithmetic
    double upR[3] = {0., 0., 0.}, nxt[3] = {0., 0., 0.}, I[3];
    // kinds 1,2: second segment fed from k-1, third (kind 2) from j-1; kinds 3,4 the other way round
    const bool secL = (kind <= 2) == (laneIsK != 0);
    if (kind == 2 || kind == 4 || (kind != 0 && !secL)) {
#pragma unroll
      for (int g = 0; g < 3; g++) upR[g] = pin[g * npl - np1];
    }
    if (q + 1 < ndir) {
      const double* pn = sp.planeIn + (int64_t)sp.pl[q + 1] * (3 * npl) + pidx;
#pragma unroll
      for (int g = 0; g < 3; g++) nxt[g] = pn[g * npl];
    }
    if (kind == 0) direction_body<FAITHFUL, EXPV, 1, true>(P, cur, upR, kap, invk, kR, I, acc, sT);
    else if (kind == 1 || kind == 3) {
      if (secL) direction_body<FAITHFUL, EXPV, 2, true>(P, cur, upR, kap, invk, kR, I, acc, sT);
      else direction_body<FAITHFUL, EXPV, 2, false>(P, cur, upR, kap, invk, kR, I, acc, sT);
    } else {
      if (secL) direction_body<FAITHFUL, EXPV, 3, true>(P, cur, upR, kap, invk, kR, I, acc, sT);
      else direction_body<FAITHFUL, EXPV, 3, false>(P, cur, upR, kap, invk, kR, I, acc, sT);
    }
    if (writer) {
#pragma unroll
      for (int g = 0; g < 3; g++) pout[g * npl] = I[g];
    }
#pragma unroll
    for (int g = 0; g < 3; g++) cur[g] = nxt[g];
  }
  if (writer) {
#pragma unroll
    for (int g = 0; g < 3; g++) {
      double* p = sp.acc + (int64_t)g * N + leaf;
      *p = sp.firstInSlot ? acc[g] : __dadd_rn(*p, acc[g]);
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
    L.invd[s] = L.d[s] > 0. ? 1.0 / L.d[s] : 0.;
  }
  L.w = weight;
  L.wn = weight / (double)L.nseg;
  L.thin = 0;
  for (int sg = 0; sg < L.nseg; sg++)
    if (len[sg] < 1e-2) L.thin = 1;
  return L;
}

static void launch_cells(int dense, int expv, bool faithful, dim3 grid, cudaStream_t s, const BatchParams& bp,
                         const double* kappa, int N) {
  dim3 block(32, 8);
  // `dense`: 0 = compiler's choice of registers (2 blocks per SM), 1 = cap for 3 blocks, 2 = cap for 4 blocks
  if (faithful) { sweep_cell_kernel<true, 0, 2><<<grid, block, 0, s>>>(bp, kappa, N); return; }
#define RTB_LAUNCH(E)                                                                       \
  if (dense == 1) sweep_cell_kernel<false, E, 3><<<grid, block, 0, s>>>(bp, kappa, N);      \
  else if (dense >= 2) sweep_cell_kernel<false, E, 4><<<grid, block, 0, s>>>(bp, kappa, N); \
  else sweep_cell_kernel<false, E, 2><<<grid, block, 0, s>>>(bp, kappa, N)
  if (expv == 1) { RTB_LAUNCH(1); } else { RTB_LAUNCH(0); }
#undef RTB_LAUNCH
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
    snprintf(buf, sizeof(buf), "%d:%a:%d:%d:%d:", n, c.boxSize, nAngularLevel, c.tune.slots, c.tune.lockstep);
    planKey = buf;
    for (const auto& d : dirs) { snprintf(buf, sizeof(buf), "%lld,", (long long)d.iray); planKey += buf; }
  }
  if (planKey != c.uniPlanKey) {
    c.uniTasks.clear();
    std::vector<RayPattern> pat;
    int64_t nseg = 0;
    int planeCursor = 0;
    for (int z = 1; z <= 24; z++) {
      std::vector<int> mine;
      for (int d = 0; d < ndir; d++)
        if (dirs[d].izone == z) mine.push_back(d);
      if (mine.empty()) continue;
      const int pieces = ((int)mine.size() + kMaxDirPerTask - 1) / kMaxDirPerTask;
      size_t o = 0;
      for (int pc = 0; pc < pieces; pc++) {
        const size_t cnt = (mine.size() - o + (pieces - pc) - 1) / (pieces - pc);  // balanced split
        UniTaskHost T;
        ZoneStrides zs = zone_strides(z, n);
        T.origin = zs.origin; T.si = zs.stride[0]; T.sj = zs.stride[1]; T.sk = zs.stride[2];
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
    {  // both plane buffers start as "boundary intensity everywhere": first-layer input and the permanent pads
      const int64_t total = (int64_t)2 * ndir * 3 * npl;
      int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)c.smCount * 16);
      fill_planes_kernel<<<blocks, 256, 0, st>>>(c.dPlanes, npl, total, uvb[0], uvb[1], uvb[2]);
    }
    // fill one task's parameters for one layer; returns the number of directions taken by this pass
    auto fill = [&](StepParams& sp, const UniTaskHost& T, int step, int pass, int first) -> int {
      sp.acc = c.dAcc + (size_t)T.slot * 3 * N;
      sp.sj = (int32_t)T.sj; sp.sk = (int32_t)T.sk;
      sp.laneIsK = T.laneIsK; sp.n = n;
      sp.planeIn = ((step & 1) ? planeA : planeB) + (size_t)T.planeFirst * 3 * npl;
      sp.planeOut = ((step & 1) ? planeB : planeA) + (size_t)T.planeFirst * 3 * npl;
      sp.origin = (int32_t)(T.origin + step * T.si);
      int cnt = 0;
      for (int q = 0; q < T.ndir; q++) {
        const LayerSeg& L = T.seg[(size_t)step * kMaxDirPerTask + q];
        const bool thin = !faithful && L.thin;
        if (thin != (pass == 1)) continue;
        sp.P[cnt] = L;
        sp.pl[cnt] = q;
        cnt++;
      }
      sp.ndir = cnt;
      sp.firstInSlot = first;
      return cnt;
    };
    // A layer whose pattern has a very short segment (a corner clip, len < 1e-2 cell) is evaluated with the
    // reference's own operation sequence even in FAST mode: there tau is tiny and the rounding noise of the
    // reference's (Iin-Iout)/log(Iin/Iout), ~1.1e-16/tau, would otherwise show up as a parity difference.
    // ~1.7% of the (direction, layer) pairs; they go to a second launch (pass 1) of the FAITHFUL kernel.
    static thread_local BatchParams bp;
    if (lockstep) {
      // all tasks of a batch advance one layer per launch; each task of a batch has its own accumulator slot
      for (int base = 0; base < ntask; base += slots) {
        const int nb = std::min(slots, ntask - base);
        std::vector<int> firstFlag(nb);
        for (int z = 0; z < nb; z++) firstFlag[z] = c.uniTasks[base + z].firstInSlot;
        for (int step = 0; step < n; step++) {
          std::vector<int> fastDone(nb, 0);
          for (int pass = 0; pass < 2; pass++) {
            int cntZ = 0;
            for (int z = 0; z < nb; z++) {
              const UniTaskHost& T = c.uniTasks[base + z];
              const int first = fastDone[z] ? 0 : firstFlag[z];
              if (fill(bp.t[cntZ], T, step, pass, first) > 0) { cntZ++; fastDone[z] = 1; }
            }
            if (cntZ == 0) continue;
            dim3 gz = grid;
            gz.z = cntZ;
            launch_cells(c.tune.minBlocks, c.tune.expVariant, faithful || pass == 1, gz, st, bp, c.dKappa, (int)N);
            nLaunched++;
          }
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
          int first = T.firstInSlot;
          for (int pass = 0; pass < 2; pass++) {
            if (fill(bp.t[0], T, step, pass, first) == 0) continue;
            first = 0;
            launch_cells(c.tune.minBlocks, c.tune.expVariant, faithful || pass == 1, grid, cs, bp, c.dKappa, (int)N);
            nLaunched++;
          }
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
             (int)faithful * 64 + c.tune.minBlocks * 4 + c.tune.expVariant + 1000 * c.tune.lockstep,
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
    int blocks = std::min<int64_t>((3 * N + 255) / 256, (int64_t)c.smCount * 16);
    merge_slots_kernel<<<blocks, 256, 0, s>>>(c.dAcc, slots, 3 * N, dJout);
  }
  RTB_CUDA(cudaGetLastError());
  if (nLaunched > 1) c.uniLaunches = nLaunched;  // a replayed graph issues what was captured
  c.lastSweepLaunches = c.uniLaunches;
  c.lastLaunches = c.uniLaunches + 2;  // + compute_opacities + merge
  return RTB200_OK;
}

}  // namespace rtb
