// micro-benchmark: FP64 FMA throughput and dependent-issue latency on the current GPU
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void fma_kernel(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
void run(int blocks, int threads, int iters) {
  double* out;
  cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  fma_kernel<ILP><<<blocks, threads>>>(out, 10, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  fma_kernel<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fmas = (double)blocks * threads * iters * ILP;
  printf("ILP=%d blocks=%d threads=%d: %.3f ms  %.2f TFMA/s (%.2f TFLOPS)  per-SM fma/clk@1.965GHz=%.1f\n", ILP, blocks,
         threads, ms, fmas / ms * 1e-9, 2 * fmas / ms * 1e-9, fmas / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
}
int main() {
  int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
  printf("%s SMs=%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  run<1>(148 * 8, 256, 20000);
  run<2>(148 * 8, 256, 20000);
  run<4>(148 * 8, 256, 20000);
  run<8>(148 * 8, 256, 10000);
  run<1>(148, 32, 100000);   // one warp per SM: dependent-issue latency
  run<1>(148, 128, 100000);  // one warp per SMSP
  run<4>(148, 128, 100000);
  run<8>(148 * 4, 512, 10000);
  return 0;
}
