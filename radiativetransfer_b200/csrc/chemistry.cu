// Ionisation equilibrium per leaf on the GPU: replaces solveRateEquations (equiSources.f90:3459-3677), the consumer of
// the transport results.  Fusing it behind the sweep / ray casting keeps HI, HeI, HeII, the six
// rate fields and Jmean1..3 on the device across the reference's outer transport <-> chemistry iteration.
//
// One thread per leaf; everything is IEEE +,-,*,/ in the reference's association order (no FMA contraction), the only
// transcendental -- log(tgas), fixed during the run (the driver holds the temperature constant, :3671-3673) -- is
// taken on the host with libm when the temperature is set.  The result is therefore bit-identical to the CPU
// restatement.  The rate-coefficient tables k1a..k6a (calc_rates.f, out of scope) are inputs.
#include <algorithm>
#include <cmath>

#include "rtb200_internal.h"

namespace rtb {

namespace {

__device__ __forceinline__ double M(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double A(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double S(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double D(double a, double b) { return __ddiv_rn(a, b); }

struct ChemParams {
  const int8_t* level;
  const double *rho, *logT;
  double *HI, *HeI, *HeII;
  const double* rates;   // [6][N] or null
  const double* J;       // [3][N] or null (uniform background)
  const double* k;       // [6][nratec]
  double ksi[6], uniform[4];
  double cellSize[32];   // physicalBoxSize / (float(2**level) * float(nx))
  double logtem0, logtem9, dlogtem, psi, mh, mhe, fourPi;
  int64_t N;             // leaves of the grid
  int64_t first, count;  // leaves [first, first + count) are solved (a rank's slab; the whole grid by default)
  int64_t rStride, jStride;  // field strides of `rates` / `J`; their element (f, leaf c) sits at f * stride + c - first
  int nratec;
  unsigned long long* maxChangeBits;
  int32_t* err;
};

__device__ __forceinline__ bool opposite(double a, double b) { return ((a > 0.) && (b < 0.)) || ((a < 0.) && (b > 0.)); }

__global__ void __launch_bounds__(128) chemistry_kernel(const __grid_constant__ ChemParams P) {
  const int64_t local = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (local >= P.count) return;
  const int64_t c = P.first + local;
  const double rho = P.rho[c];
  const double nh = D(M(P.psi, rho), P.mh);
  const double onemPsi = S(1., P.psi);
  const double nhe = D(M(onemPsi, rho), P.mhe);
  const double cHI = P.HI[c];
  double cHeI = P.HeI[c], cHeII = P.HeII[c];
  double HI = fmin(cHI, nh);
  double HII = S(nh, cHI);
  double HeI = cHeI, HeII = cHeII;
  double HeIII = S(S(nhe, cHeI), cHeII);
  if (HeIII < 0.) {
    cHeII = S(nhe, cHeI);
    HeIII = 0.;
    if (HeII < 0.) { cHeI = nhe; cHeII = 0.; HeII = 0.; HeIII = 0.; }
  }
  const double pcs = P.cellSize[P.level[c]];
  const double vol = M(M(pcs, pcs), pcs);
  double krate24 = 0., krate25 = 0., krate26 = 0.;
  if (P.rates) {
    if (HI > 0.) krate24 = D(P.rates[local], M(vol, HI));
    if (HeII > 0.) krate25 = D(P.rates[P.rStride + local], M(vol, HeII));
    if (HeI > 0.) krate26 = D(P.rates[2 * P.rStride + local], M(vol, HeI));
  }
  krate24 = fmax(krate24, 0.); krate25 = fmax(krate25, 0.); krate26 = fmax(krate26, 0.);
  if (P.J) {
    const double t1 = M(P.fourPi, P.J[local]), t2 = M(P.fourPi, P.J[P.jStride + local]),
                 t3 = M(P.fourPi, P.J[2 * P.jStride + local]);
    krate24 = A(A(A(krate24, M(t1, P.ksi[0])), M(t2, P.ksi[1])), M(t3, P.ksi[2]));
    krate25 = A(krate25, M(t3, P.ksi[3]));
    krate26 = A(A(krate26, M(t2, P.ksi[4])), M(t3, P.ksi[5]));
  } else {
    const double mfp = D(1., A(A(M(HI, (double)6.3e-18f), M(HeI, (double)7.42e-18f)), M(HeII, (double)1.58e-18f)));
    if (mfp >= P.uniform[3]) {
      krate24 = A(krate24, P.uniform[0]);
      krate25 = A(krate25, P.uniform[1]);
      krate26 = A(krate26, P.uniform[2]);
    }
  }
  double logtem = fmin(fmax(P.logT[c], P.logtem0), P.logtem9);
  const int indixe = min(P.nratec - 1, max(1, (int)D(S(logtem, P.logtem0), P.dlogtem) + 1));
  const double t1 = A(P.logtem0, M((double)(indixe - 1), P.dlogtem)), t2 = A(P.logtem0, M((double)indixe, P.dlogtem));
  const double tdef = S(t2, t1), dt = S(logtem, t1);
  double kk[6];
#pragma unroll
  for (int i = 0; i < 6; i++) {
    const double* t = P.k + (size_t)i * P.nratec;
    kk[i] = A(t[indixe - 1], D(M(dt, S(t[indixe], t[indixe - 1])), tdef));
  }
  const double k1 = kk[0], k2 = kk[1], k3 = kk[2], k4 = kk[3], k5 = kk[4], k6 = kk[5];
  const double twoNhe = M(2., nhe);
  auto heI = [&](double d) {
    const double q = D(A(M(k3, d), krate26), M(k4, d));
    const double num = S(S(d, D(nh, A(1., D(M(k2, d), A(M(k1, d), krate24))))), twoNhe);
    const double den = S(S(q, 2.), D(M(2., A(M(k3, d), krate26)), M(k4, d)));
    return D(num, den);
  };
  auto resid = [&](double h, double d) {
    const double x = D(M(h, A(M(k3, d), krate26)), M(k4, d));
    const double a1 = M(M(k3, h), d);
    const double a2 = M(M(k6, S(S(nhe, h), x)), d);
    const double a3 = M(krate26, h);
    const double a4 = M(x, A(A(M(k4, d), M(k5, d)), krate25));
    return S(A(A(a1, a2), a3), a4);
  };
  double de1 = (double)1.e-30f, de = de1;
  HeI = heI(de);
  double res1 = resid(HeI, de);
  double de2 = A(nh, twoNhe);
  de = de2;
  HeI = heI(de);
  double HeIprev = -1.;
  int guard = 0;
  while (D(fabs(S(HeI, HeIprev)), nhe) > 1.e-10) {
    HeIprev = HeI;
    de = M(0.5, A(de1, de2));
    HeI = heI(de);
    const double res = resid(HeI, de);
    if (opposite(res, res1)) de2 = de;
    else { de1 = de; res1 = res; }
    if (++guard > 100000) { atomicExch(P.err, RTB200_ERR_CHEMISTRY); return; }
  }
  HeII = D(M(HeI, A(M(k3, de), krate26)), M(k4, de));
  HII = D(nh, A(1., D(M(k2, de), A(M(k1, de), krate24))));
  HI = D(M(M(k2, HII), de), A(M(k1, de), krate24));
  const double fH = D(HI, nh), fHe = D(HeI, nhe);
  if (!(fH >= 0. && fH <= 1.) || !(fHe >= 0. && fHe <= 1.)) { atomicExch(P.err, RTB200_ERR_CHEMISTRY); return; }
  const double c1 = D(M(fabs(S(HI, cHI)), P.mh), M(P.psi, rho));
  const double c2 = D(M(fabs(S(HeI, cHeI)), P.mhe), M(onemPsi, rho));
  const double c3 = D(M(fabs(S(HeII, cHeII)), P.mhe), M(onemPsi, rho));
  const double worst = fmax(c1, fmax(c2, c3));
  atomicMax(P.maxChangeBits, (unsigned long long)__double_as_longlong(worst));   // non-negative doubles order like integers
  P.HI[c] = HI; P.HeI[c] = HeI; P.HeII[c] = HeII;
}

}  // namespace

int chemistry_set_tables(Context& c, int nratec, double logtem0, double logtem9, double dlogtem, const double* const k[6]) {
  if (nratec < 2 || !(dlogtem > 0.)) return RTB200_ERR_ARG;
  for (int i = 0; i < 6; i++)
    if (!k[i]) return RTB200_ERR_ARG;
  RTB_CUDA(cudaSetDevice(c.device));
  if (c.dChemK) { cudaFree(c.dChemK); c.dChemK = nullptr; }
  RTB_CUDA(cudaMalloc((void**)&c.dChemK, (size_t)6 * nratec * sizeof(double)));
  for (int i = 0; i < 6; i++)
    RTB_CUDA(cudaMemcpy(c.dChemK + (size_t)i * nratec, k[i], (size_t)nratec * sizeof(double), cudaMemcpyHostToDevice));
  c.chemNratec = nratec; c.chemLogtem0 = logtem0; c.chemLogtem9 = logtem9; c.chemDlogtem = dlogtem;
  return RTB200_OK;
}

int chemistry_set_temperature(Context& c, const double* tgas) {
  if (!tgas || c.nleaf == 0) return RTB200_ERR_ARG;
  RTB_CUDA(cudaSetDevice(c.device));
  std::vector<double> lt((size_t)c.nleaf);
  for (int64_t i = 0; i < c.nleaf; i++) lt[(size_t)i] = std::log(tgas[i]);   // libm, as the reference (:3568)
  if (!c.dLogT) RTB_CUDA(cudaMalloc((void**)&c.dLogT, (size_t)c.nleaf * sizeof(double)));
  RTB_CUDA(cudaMemcpy(c.dLogT, lt.data(), (size_t)c.nleaf * sizeof(double), cudaMemcpyHostToDevice));
  return RTB200_OK;
}

static int chemistry_launch(Context& c, int64_t first, int64_t count, const double* dRates, int64_t rStride,
                            const double* dJ, int64_t jStride, const double* ksi, const double* uniform,
                            double* maxChange, cudaStream_t s);

int chemistry_run(Context& c, const double* dRates, const double* dJ, const double* ksi, const double* uniform,
                  double* maxChange, cudaStream_t s) {
  return chemistry_launch(c, 0, c.nleaf, dRates, c.nleaf, dJ, c.nleaf, ksi, uniform, maxChange, s);
}

int chemistry_run_slab(Context& c, int64_t first, int64_t count, const double* dRates, int64_t rStride, const double* dJ,
                       int64_t jStride, const double* ksi, const double* uniform, cudaStream_t s) {
  if (first < 0 || count < 0 || first + count > c.nleaf) return RTB200_ERR_ARG;
  if (count == 0) return RTB200_OK;
  return chemistry_launch(c, first, count, dRates, rStride, dJ, jStride, ksi, uniform, nullptr, s);
}

static int chemistry_launch(Context& c, int64_t first, int64_t count, const double* dRates, int64_t rStride,
                            const double* dJ, int64_t jStride, const double* ksi, const double* uniform,
                            double* maxChange, cudaStream_t s) {
  if (c.nleaf == 0 || !c.dRho || !c.dChemK || !c.dLogT) return RTB200_ERR_ARG;
  if (dJ && !ksi) return RTB200_ERR_ARG;
  if (!dJ && !uniform) return RTB200_ERR_ARG;
  RTB_CUDA(cudaSetDevice(c.device));
  ChemParams P{};
  P.first = first; P.count = count; P.rStride = rStride; P.jStride = jStride;
  P.level = c.dLevel; P.rho = c.dRho; P.logT = c.dLogT; P.HI = c.dHI; P.HeI = c.dHeI; P.HeII = c.dHeII;
  P.rates = dRates; P.J = dJ; P.k = c.dChemK;
  for (int i = 0; i < 6; i++) P.ksi[i] = ksi ? ksi[i] : 0.;
  for (int i = 0; i < 4; i++) P.uniform[i] = uniform ? uniform[i] : 0.;
  for (int l = 0; l < 32; l++) P.cellSize[l] = l < 31 ? c.boxSize / ((double)(float)(1u << l) * (double)(float)c.nx) : 0.;
  P.logtem0 = c.chemLogtem0; P.logtem9 = c.chemLogtem9; P.dlogtem = c.chemDlogtem;
  P.psi = (double)0.76f; P.mh = (double)1.6726231e-24f;
  P.mhe = 2.0 * ((double)1.6726231e-24f + (double)1.67492728e-24f);
  P.fourPi = 4. * kPi;
  P.N = c.nleaf; P.nratec = c.chemNratec;
  unsigned long long* dMax = (unsigned long long*)(c.dErr + 8);   // the error block holds 64 bytes: [0] status, [8..9] max change
  RTB_CUDA(cudaMemsetAsync(dMax, 0, 8, s));
  P.maxChangeBits = dMax; P.err = c.dErr;
  chemistry_kernel<<<(unsigned)((count + 127) / 128), 128, 0, s>>>(P);
  RTB_CUDA(cudaGetLastError());
  if (maxChange) {
    unsigned long long bits = 0;
    RTB_CUDA(cudaMemcpyAsync(&bits, dMax, 8, cudaMemcpyDeviceToHost, s));
    RTB_CUDA(cudaStreamSynchronize(s));
    double v;
    memcpy(&v, &bits, 8);
    *maxChange = v;
    int32_t err = 0;
    RTB_CUDA(cudaMemcpy(&err, c.dErr, sizeof(err), cudaMemcpyDeviceToHost));
    if (err) { cudaMemset(c.dErr, 0, 64); return err; }
  }
  return RTB200_OK;
}

// ---------------------------------------------------------------------------------------------------------
// computeMass (equiSources.f90:4369-4393): neutral and total hydrogen mass of the grid in solar masses, from the
// absorber densities the context holds now (i.e. after the last chemistry step).  Each leaf's two terms follow the
// reference's operation order; the reference adds them serially in leaf order, here the sum is a fixed two-stage
// tree (1024 blocks of 256 partial sums, then one block), so it is reproducible run to run and differs from the
// serial sum only by rounding (~1e-16 sqrt(N) relative).
// ---------------------------------------------------------------------------------------------------------
struct MassParams {
  const int8_t* level;
  const double* HI;
  const double* rho;
  double cube[32];   // (physicalBoxSize / (float(2**level) * float(nx)))**3
  double mh, msun, psi;
  int64_t N;
};
constexpr int kMassBlocks = 1024;

__device__ __forceinline__ void block_sum2(double& a, double& b, double* sh) {
  // fixed-order tree over the 256 threads of the block
  const int t = threadIdx.x;
  sh[t] = a; sh[256 + t] = b;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (t < w) {
      sh[t] = __dadd_rn(sh[t], sh[t + w]);
      sh[256 + t] = __dadd_rn(sh[256 + t], sh[256 + t + w]);
    }
    __syncthreads();
  }
  a = sh[0]; b = sh[256];
}

__global__ void __launch_bounds__(256) mass_partial_kernel(const __grid_constant__ MassParams P, double* __restrict__ part) {
  __shared__ double sh[512];
  double neutral = 0., total = 0.;
  // a block owns one contiguous slab of leaves, a thread every 256th leaf of it: the order is fixed by (N) alone
  const int64_t per = (P.N + kMassBlocks - 1) / kMassBlocks;
  const int64_t lo = blockIdx.x * per, hi = lo + per < P.N ? lo + per : P.N;
  for (int64_t i = lo + threadIdx.x; i < hi; i += 256) {
    const double cube = P.cube[P.level[i]];
    neutral = __dadd_rn(neutral, __ddiv_rn(__dmul_rn(__dmul_rn(P.HI[i], P.mh), cube), P.msun));
    total = __dadd_rn(total, __ddiv_rn(__dmul_rn(__dmul_rn(P.psi, P.rho[i]), cube), P.msun));
  }
  block_sum2(neutral, total, sh);
  if (threadIdx.x == 0) { part[blockIdx.x] = neutral; part[kMassBlocks + blockIdx.x] = total; }
}

__global__ void __launch_bounds__(256) mass_final_kernel(const double* __restrict__ part, double* __restrict__ out) {
  __shared__ double sh[512];
  double neutral = 0., total = 0.;
  for (int i = threadIdx.x; i < kMassBlocks; i += 256) {
    neutral = __dadd_rn(neutral, part[i]);
    total = __dadd_rn(total, part[kMassBlocks + i]);
  }
  block_sum2(neutral, total, sh);
  if (threadIdx.x == 0) { out[0] = neutral; out[1] = total; }
}

int compute_mass(Context& c, double* neutralHydrogenMass, double* totalHydrogenMass, cudaStream_t s) {
  if (c.nleaf == 0 || !c.dRho || !c.dHI || !neutralHydrogenMass || !totalHydrogenMass) return RTB200_ERR_ARG;
  RTB_CUDA(cudaSetDevice(c.device));
  MassParams P{};
  P.level = c.dLevel; P.HI = c.dHI; P.rho = c.dRho; P.N = c.nleaf;
  for (int l = 0; l < 32; l++) {
    const double size = l < 31 ? c.boxSize / ((double)(float)(1u << l) * (double)(float)c.nx) : 0.;
    P.cube[l] = size * size * size;            // x**3 -> x*x*x
  }
  P.mh = (double)1.6726231e-24f;               // definitionsModule.f90:27 (single-precision literal)
  P.msun = (double)1.98892e33f;                // :29
  P.psi = (double)0.76f;                       // :261
  if (!c.dMassPart) RTB_CUDA(cudaMalloc((void**)&c.dMassPart, (2 * kMassBlocks + 2) * sizeof(double)));
  mass_partial_kernel<<<kMassBlocks, 256, 0, s>>>(P, c.dMassPart);
  mass_final_kernel<<<1, 256, 0, s>>>(c.dMassPart, c.dMassPart + 2 * kMassBlocks);
  RTB_CUDA(cudaGetLastError());
  double out[2];
  RTB_CUDA(cudaMemcpyAsync(out, c.dMassPart + 2 * kMassBlocks, sizeof(out), cudaMemcpyDeviceToHost, s));
  RTB_CUDA(cudaStreamSynchronize(s));
  *neutralHydrogenMass = out[0];
  *totalHydrogenMass = out[1];
  return RTB200_OK;
}

}  // namespace rtb
