// Device microbenchmarks used for roofline denominators (bench.py): FP64 FMA peak of the part we run on.
#include "kernels.cuh"

namespace b200 {

// 8 independent DFMA chains per thread, all in registers: measures the FP64 pipe, nothing else
__global__ void __launch_bounds__(256) k_dfma_peak(double* __restrict__ out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0, x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0,
         x7 = x0 + 7.0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456) out[0] = s;  // never true; keeps the chains alive
}

// returns achieved TFLOP/s (2 flops per FMA), best of `reps`
double measure_dfma_tflops(cudaStream_t st, int sm_count, int reps) {
  double* d = nullptr;
  if (cudaMalloc((void**)&d, 64) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = sm_count * 8, threads = 256, iters = 4096;
  k_dfma_peak<<<blocks, threads, 0, st>>>(d, 64, 0.999999, 1e-9);  // warm-up
  double best = 0.0;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0, st);
    k_dfma_peak<<<blocks, threads, 0, st>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
    if (ms > 0.f) best = fmax(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  return cudaGetLastError() == cudaSuccess ? best : -1.0;
}

}  // namespace b200
