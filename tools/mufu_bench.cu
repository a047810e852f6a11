// Microbenchmark: MUFU (ex2 / rcp) and FFMA issue rates per SM on this GPU, alone and mixed.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); }
      if (MODE == 1) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i])); }
      if (MODE == 2) { a[i] = fmaf(a[i], 1.0001f, 0.5f); }
      if (MODE == 3) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); a[i] = fmaf(a[i], 1.0001f, 0.5f); a[i] = fmaf(a[i], 0.999f, 0.25f); a[i] = fmaf(a[i], 0.999f, 0.25f); a[i] = fmaf(a[i], 0.999f, 0.25f); }
      if (MODE == 4) { asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i])); }
    }
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int ops_per_inner, int warps) {
  float* d; cudaMalloc(&d, 148 * 1024 * 4);
  int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148, warps * 32>>>(d, 100, 1.0f);
  cudaEventRecord(e0); k<MODE><<<148, warps * 32>>>(d, iters, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double ops = (double)iters * 8 * ops_per_inner * warps * 32;  // per SM
  printf("%-28s warps=%2d  %.3f ms  %.2f thread-ops/clk/SM (at %d MHz nominal)\n", name, warps, ms, ops / (ms * 1e-3 * clk * 1e3), clk / 1000);
  cudaFree(d);
}
int main() {
  for (int w : {4, 8, 16, 32}) {
    run<0>("ex2.approx", 1, w); run<1>("rcp.approx", 1, w); run<4>("tanh.approx", 1, w); run<2>("ffma", 1, w); run<3>("1 ex2 + 4 ffma (count ex2)", 1, w);
  }
  return 0;
}
