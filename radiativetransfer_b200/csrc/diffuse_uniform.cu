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
// one layer step of up to gridDim.z zones
// ---------------------------------------------------------------------------------------------------------
template <int TY, bool FAITHFUL>
__global__ void __launch_bounds__(32 * TY)
sweep_layer_kernel(const UniTask* __restrict__ tasks, int taskBase, int step, int n, const LayerSeg* __restrict__ pats,
                   const double* __restrict__ kappa, int64_t N, double u0, double u1, double u2,
                   const double* __restrict__ planeIn, double* __restrict__ planeOut) {
  __shared__ double sm[2][3][TY][33];
  const UniTask& T = tasks[taskBase + blockIdx.z];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int a = blockIdx.x * 31 - 1 + tx;        // coordinate along the lane axis (0-based), halo at tx == 0
  const int b = blockIdx.y * (TY - 1) - 1 + ty;  // coordinate along the other axis, halo at ty == 0
  const bool inDom = a >= 0 && a < n && b >= 0 && b < n;
  const bool writer = inDom && tx >= 1 && ty >= 1;
  const int laneIsK = T.laneIsK;
  const int j = laneIsK ? b : a, k = laneIsK ? a : b;
  const int64_t nn = (int64_t)n * n;
  const int64_t pidx = (int64_t)b * n + a;
  double kap[3] = {0., 0., 0.}, invk[3] = {0., 0., 0.};
  bool kpos[3] = {false, false, false};
  int64_t leaf = 0;
  if (inDom) {
    leaf = T.origin + step * T.si + j * T.sj + k * T.sk;
#pragma unroll
    for (int g = 0; g < 3; g++) {
      kap[g] = kappa[g * N + leaf];
      kpos[g] = kap[g] > 0.;
      if (!FAITHFUL) invk[g] = 1.0 / kap[g];
    }
  }
  // upstream neighbours in shared memory: "from k-1" and "from j-1"
  const int txm = tx > 0 ? tx - 1 : 0, tym = ty > 0 ? ty - 1 : 0;
  const int kx = laneIsK ? txm : tx, ky = laneIsK ? ty : tym;  // cell (j, k-1)
  const int jx = laneIsK ? tx : txm, jy = laneIsK ? tym : ty;  // cell (j-1, k)
  const bool kEdge = (k == 0), jEdge = (j == 0);
  const double uvb[3] = {u0, u1, u2};
  double acc[3] = {0., 0., 0.};
  int ring = 0;
  for (int q = 0; q < T.ndir; q++) {
    const int d = T.dir[q];
    const LayerSeg P = pats[(int64_t)d * n + step];
    const double* pin = planeIn + (int64_t)d * 3 * nn;
    double I[3], Js[3], J2[3] = {0., 0., 0.}, J3[3] = {0., 0., 0.};
    if (inDom) {
#pragma unroll
      for (int g = 0; g < 3; g++) {
        double Iin = (step == 0) ? uvb[g] : pin[g * nn + pidx];
        SegResult r = segment_update<FAITHFUL>(Iin, kap[g], P.d[0], invk[g] * P.invd[0], kpos[g]);
        I[g] = r.Iout;
        Js[g] = r.J;
      }
    } else {
      I[0] = I[1] = I[2] = 0.; Js[0] = Js[1] = Js[2] = 0.;
    }
    if (P.kind != 0) {  // uniform across the block
      const bool secondFromK = P.kind <= 2;
#pragma unroll
      for (int g = 0; g < 3; g++) sm[ring][g][ty][tx] = I[g];
      __syncthreads();
      if (inDom) {
        const bool edge = secondFromK ? kEdge : jEdge;
        const int sx = secondFromK ? kx : jx, sy = secondFromK ? ky : jy;
#pragma unroll
        for (int g = 0; g < 3; g++) {
          double Iin = edge ? uvb[g] : sm[ring][g][sy][sx];
          SegResult r = segment_update<FAITHFUL>(Iin, kap[g], P.d[1], invk[g] * P.invd[1], kpos[g]);
          I[g] = r.Iout;
          J2[g] = r.J;
        }
      }
      ring ^= 1;
      if (P.kind == 2 || P.kind == 4) {
#pragma unroll
        for (int g = 0; g < 3; g++) sm[ring][g][ty][tx] = I[g];
        __syncthreads();
        if (inDom) {
          const bool edge = secondFromK ? jEdge : kEdge;
          const int sx = secondFromK ? jx : kx, sy = secondFromK ? jy : ky;
#pragma unroll
          for (int g = 0; g < 3; g++) {
            double Iin = edge ? uvb[g] : sm[ring][g][sy][sx];
            SegResult r = segment_update<FAITHFUL>(Iin, kap[g], P.d[2], invk[g] * P.invd[2], kpos[g]);
            I[g] = r.Iout;
            J3[g] = r.J;
          }
        }
        ring ^= 1;
      }
    }
    if (writer) {
      double* pout = planeOut + (int64_t)d * 3 * nn;
#pragma unroll
      for (int g = 0; g < 3; g++) {
        pout[g * nn + pidx] = I[g];
        // the reference sums the segments in the order xy, xz, yz (transportRoutinesModule.f90:698,818,941)
        const bool yzSecond = P.kind <= 2;  // kinds 1,2: second segment is the yz ray, third the xz ray
        double Jxz = yzSecond ? J3[g] : J2[g];
        double Jyz = yzSecond ? J2[g] : J3[g];
        double sum = Js[g];
        if (P.kind >= 2) sum = __dadd_rn(sum, Jxz);                  // xz ray active: kinds 2, 3, 4
        if (P.kind != 0 && P.kind != 3) sum = __dadd_rn(sum, Jyz);   // yz ray active: kinds 1, 2, 4
        if (FAITHFUL) acc[g] = __dadd_rn(acc[g], __dmul_rn(__ddiv_rn(sum, (double)P.nseg), P.w));
        else acc[g] = fma(sum, P.wn, acc[g]);
      }
    }
  }
  if (writer) {
#pragma unroll
    for (int g = 0; g < 3; g++) {
      double* p = T.acc + g * N + leaf;
      *p = T.firstInSlot ? acc[g] : __dadd_rn(*p, acc[g]);
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
  return L;
}

template <int TY>
static void launch_layer(bool faithful, dim3 grid, cudaStream_t s, const UniTask* tasks, int taskBase, int step, int n,
                         const LayerSeg* pats, const double* kappa, int64_t N, const double* uvb, const double* pin,
                         double* pout) {
  dim3 block(32, TY);
  if (faithful)
    sweep_layer_kernel<TY, true><<<grid, block, 0, s>>>(tasks, taskBase, step, n, pats, kappa, N, uvb[0], uvb[1],
                                                        uvb[2], pin, pout);
  else
    sweep_layer_kernel<TY, false><<<grid, block, 0, s>>>(tasks, taskBase, step, n, pats, kappa, N, uvb[0], uvb[1],
                                                         uvb[2], pin, pout);
}

int diffuse_uniform(Context& c, int nAngularLevel, const double* uvb, const std::vector<Direction>& dirs,
                    double* dJout, cudaStream_t s, int64_t* nsegOut) {
  const int n = c.nx;
  const int64_t N = c.nleaf, nn = (int64_t)n * n;
  const int ndir = (int)dirs.size();
  const int64_t nraysTotal = 12LL << (2 * (nAngularLevel - 1));
  const double weight = (double)(1.f / (float)nraysTotal);  // equiSources.f90:1386 (single-precision division)
  const double cellSize = c.boxSize / (double)n;             // equiSources.f90:1570

  // ---- plan: pattern tables, zone tasks, slot assignment.  Depends only on the grid size and the direction
  //      list, so it is cached across the outer transport<->chemistry iterations. ----
  std::string planKey;
  {
    char buf[128];
    snprintf(buf, sizeof(buf), "%d:%a:%d:%d:%d:", n, c.boxSize, nAngularLevel, c.tune.slots, (int)c.tune.l2BudgetMB);
    planKey = buf;
    for (const auto& d : dirs) { snprintf(buf, sizeof(buf), "%lld,", (long long)d.iray); planKey += buf; }
  }
  if (planKey != c.uniPlanKey) {
    std::vector<LayerSeg> hp((size_t)ndir * n);
    std::vector<RayPattern> pat;
    int64_t nseg = 0;
    for (int d = 0; d < ndir; d++) {
      if (dirs[d].status) return dirs[d].status;
      layer_patterns_level0(dirs[d].phi, dirs[d].theta, n, pat);
      for (int i = 0; i < n; i++) {
        if (pat[i].status) return pat[i].status;
        hp[(size_t)d * n + i] = make_layer_seg(pat[i], cellSize, weight);
        nseg += (int64_t)hp[(size_t)d * n + i].nseg * nn;
      }
    }
    // tasks: directions grouped by zone, split into chunks of kMaxDirPerTask
    std::vector<UniTask> tasks;
    for (int z = 1; z <= 24; z++) {
      std::vector<int> mine;
      for (int d = 0; d < ndir; d++)
        if (dirs[d].izone == z) mine.push_back(d);
      for (size_t o = 0; o < mine.size(); o += kMaxDirPerTask) {
        UniTask T;
        std::memset(&T, 0, sizeof(T));
        ZoneStrides zs = zone_strides(z, n);
        T.origin = zs.origin; T.si = zs.stride[0]; T.sj = zs.stride[1]; T.sk = zs.stride[2];
        T.ndir = (int)std::min<size_t>(kMaxDirPerTask, mine.size() - o);
        for (int q = 0; q < T.ndir; q++) T.dir[q] = mine[o + q];
        int64_t aj = T.sj < 0 ? -T.sj : T.sj, ak = T.sk < 0 ? -T.sk : T.sk;
        T.laneIsK = ak <= aj;
        tasks.push_back(T);
      }
    }
    const int ntask = (int)tasks.size();
    // heaviest first so that the tasks sharing a launch have similar cost
    std::stable_sort(tasks.begin(), tasks.end(), [](const UniTask& x, const UniTask& y) { return x.ndir > y.ndir; });
    // zones in flight: the planes of the in-flight zones (ping + pong) should fit the L2 budget
    int slots = c.tune.slots;
    if (slots <= 0) {
      double perTask = 2.0 * 8 * 3 * nn * 8.0;  // ~8 directions per zone at nAngularLevel 3
      slots = (int)(c.tune.l2BudgetMB * 1048576.0 / perTask);
      slots = std::max(2, std::min(slots, 24));
    }
    slots = std::max(1, std::min(slots, ntask));
    if (int st = ensure_buffer((void**)&c.dAcc, &c.accBytes, (size_t)slots * 3 * N * sizeof(double))) return st;
    if (int st = ensure_buffer((void**)&c.dPlanes, &c.planeBytes, (size_t)2 * std::max(ndir, 1) * 3 * nn * sizeof(double))) return st;
    if (int st = ensure_buffer(&c.dTasks, &c.taskBytes, (size_t)std::max(ntask, 1) * sizeof(UniTask))) return st;
    if (int st = ensure_buffer(&c.dPats, &c.patBytes, std::max<size_t>(hp.size(), 1) * sizeof(LayerSeg))) return st;
    std::vector<int> seen(slots, 0);
    for (int t = 0; t < ntask; t++) {
      int slot = t % slots;
      tasks[t].acc = c.dAcc + (size_t)slot * 3 * N;
      tasks[t].firstInSlot = !seen[slot];
      seen[slot] = 1;
    }
    if (ntask) {
      RTB_CUDA(cudaMemcpyAsync(c.dTasks, tasks.data(), (size_t)ntask * sizeof(UniTask), cudaMemcpyHostToDevice, s));
      RTB_CUDA(cudaMemcpyAsync(c.dPats, hp.data(), hp.size() * sizeof(LayerSeg), cudaMemcpyHostToDevice, s));
      RTB_CUDA(cudaStreamSynchronize(s));  // pageable sources go out of scope below
    }
    c.uniPlanKey = planKey;
    c.uniNtask = ntask;
    c.uniSlots = slots;
    c.uniNseg = nseg;
    if (c.graphExec) { cudaGraphExecDestroy(c.graphExec); c.graphExec = nullptr; }
  }
  const int ntask = c.uniNtask, slots = c.uniSlots;
  if (nsegOut) *nsegOut = c.uniNseg;
  if (ntask == 0) {
    RTB_CUDA(cudaMemsetAsync(dJout, 0, 3 * N * sizeof(double), s));
    return RTB200_OK;
  }

  const int TY = c.tune.tileY == 8 ? 8 : 16;
  dim3 grid((n + 30) / 31, (n + TY - 2) / (TY - 1), 1);
  double* planeA = c.dPlanes;
  double* planeB = c.dPlanes + (size_t)ndir * 3 * nn;
  const bool faithful = c.mathMode == RTB200_MATH_FAITHFUL;
  int64_t launches = 0;

  auto issue = [&](cudaStream_t st) -> int {
    for (int base = 0; base < ntask; base += slots) {
      grid.z = std::min(slots, ntask - base);
      for (int step = 0; step < n; step++) {
        const double* pin = (step & 1) ? planeA : planeB;
        double* pout = (step & 1) ? planeB : planeA;
        switch (TY) {
          case 8: launch_layer<8>(faithful, grid, st, (const UniTask*)c.dTasks, base, step, n, (const LayerSeg*)c.dPats, c.dKappa, N, uvb, pin, pout); break;
          default: launch_layer<16>(faithful, grid, st, (const UniTask*)c.dTasks, base, step, n, (const LayerSeg*)c.dPats, c.dKappa, N, uvb, pin, pout); break;
        }
        launches++;
      }
    }
    int blocks = std::min<int64_t>((3 * N + 255) / 256, (int64_t)c.smCount * 16);
    merge_slots_kernel<<<blocks, 256, 0, st>>>(c.dAcc, slots, 3 * N, dJout);
    launches++;
    return RTB200_OK;
  };

  if (c.tune.useGraph) {
    // The launch sequence depends only on (n, ntask, slots, TY, mode, buffers): capture once, replay afterwards.
    char key[256];
    snprintf(key, sizeof(key), "u:%d:%d:%d:%d:%d:%p:%p:%p:%p:%p:%g:%g:%g", n, ntask, slots, TY, (int)faithful,
             (void*)c.dAcc, (void*)c.dPlanes, c.dTasks, c.dPats, (void*)dJout, uvb[0], uvb[1], uvb[2]);
    if (!c.graphExec || c.graphKey != key) {
      if (c.graphExec) { cudaGraphExecDestroy(c.graphExec); c.graphExec = nullptr; }
      cudaGraph_t graph;
      RTB_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
      issue(s);
      RTB_CUDA(cudaStreamEndCapture(s, &graph));
      RTB_CUDA(cudaGraphInstantiate(&c.graphExec, graph, 0));
      cudaGraphDestroy(graph);
      c.graphKey = key;
    } else {
      launches = (int64_t)((ntask + slots - 1) / slots) * n + 1;
    }
    RTB_CUDA(cudaGraphLaunch(c.graphExec, s));
  } else {
    issue(s);
  }
  RTB_CUDA(cudaGetLastError());
  c.lastLaunches = launches + 1;  // + compute_opacities
  return RTB200_OK;
}

}  // namespace rtb
