// FP64 pipe microbenchmark for sm_100a: dependent-issue latency and per-SM throughput of the instructions the pair
// sweep is made of.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_microbench fp64_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__device__ __forceinline__ double op(double a, double b, double c) {
  double r;
  if (OP == 0) asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(r) : "d"(a), "d"(b), "d"(c));
  else if (OP == 1) asm volatile("add.rn.f64 %0, %1, %2;" : "=d"(r) : "d"(a), "d"(c));
  else if (OP == 2) asm volatile("mul.rn.f64 %0, %1, %2;" : "=d"(r) : "d"(a), "d"(b));
  else if (OP == 3) asm volatile("add.rm.f64 %0, %1, %2;" : "=d"(r) : "d"(a), "d"(c));
  else if (OP == 4) { int lo, hi; asm volatile("{.reg .pred p; setp.gt.f64 p, %2, %3; selp.b32 %0, 1, 0, p; mov.b32 %1, 0;}" : "=r"(lo), "=r"(hi) : "d"(a), "d"(c)); r = a + (double)lo; }
  else { int h; asm volatile("{.reg .b32 lo; mov.b64 {lo, %0}, %1;}" : "=r"(h) : "d"(a)); float f; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(f) : "f"(__int_as_float(h))); r = a + (double)f; }
  return r;
}

// latency: one warp, one dependent chain
template <int OP>
__global__ void k_lat(double* out, long long* cyc, int n) {
  double a = threadIdx.x * 1e-3 + 1.0, b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) a = op<OP>(a, b, c);
  }
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
// throughput: many warps, 8 independent chains each
template <int OP>
__global__ void k_thr(double* out, int n) {
  double a[8];
  for (int u = 0; u < 8; ++u) a[u] = threadIdx.x * 1e-3 + u;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] = op<OP>(a[u], b, c);
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] = op<OP>(a[u], b, c);
  }
  double s = 0;
  for (int u = 0; u < 8; ++u) s += a[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, double* d_out, long long* d_cyc, int sms) {
  const int n = 4096;
  k_lat<OP><<<1, 32>>>(d_out, d_cyc, n);
  k_lat<OP><<<1, 32>>>(d_out, d_cyc, n);
  long long cyc;
  cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = sms * 2, threads = 512, m = 8192;
  k_thr<OP><<<blocks, threads>>>(d_out, m);
  cudaEventRecord(e0);
  k_thr<OP><<<blocks, threads>>>(d_out, m);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double ops = (double)blocks * threads * m * 16.0;
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("%-12s dependent latency %.2f cycles; throughput %.1f lane-ops/clk/SM (at %d MHz nominal), %.2f Tops/s\n", name,
         (double)cyc / (n * 16.0), ops / (ms * 1e-3) / (clk_khz * 1e3) / sms, clk_khz / 1000, ops / (ms * 1e-3) / 1e12);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* d_out;
  long long* d_cyc;
  cudaMalloc(&d_out, sizeof(double) * sms * 2 * 512);
  cudaMalloc(&d_cyc, sizeof(long long));
  run<0>("DFMA", d_out, d_cyc, sms);
  run<1>("DADD", d_out, d_cyc, sms);
  run<2>("DMUL", d_out, d_cyc, sms);
  run<3>("DADD.RM", d_out, d_cyc, sms);
  run<4>("DSETP+I2F", d_out, d_cyc, sms);
  run<5>("MUFU+F2F", d_out, d_cyc, sms);
  return 0;
}
