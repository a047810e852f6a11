// Throughput of tcgen05.mma kind::f16 (M = 128, K = 16 per instruction) for the shapes and operand sources the kernels of
// this repo issue: clocks per instruction for a long stream from ONE thread, operands in shared memory (SS) or A in tensor
// memory (TS), accumulating into one TMEM tile or alternating between two.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../neurallaplacecontrol_b200/csrc -o mma_rate mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_umma.cuh"
using namespace nlc::umma;

template <int N, bool kTS, int kAlt>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ alignas(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (64 + 48) * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 64 * 1024);
    const uint32_t idesc = idesc_f16_f32(128, N);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t off = ks * 256;
        const uint32_t d = tmem + ((kAlt > 1) ? ((ks % kAlt) * 256) : 0);
        if (kTS) mma_f16_ts(d, tmem + 448 + 8 * ks, smem_desc(b + off, 128, 1024), idesc, 1u);
        else mma_f16_ss(d, smem_desc(a + off, 128, 1024), smem_desc(b + off, 128, 1024), idesc, 1u);
      }
    }
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int N, bool kTS, int kAlt>
void run(const char* name) {
  long long* d;
  cudaMalloc(&d, 148 * 8);
  const int iters = 2000, smem = (64 + 48) * 1024;
  cudaFuncSetAttribute(k<N, kTS, kAlt>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<N, kTS, kAlt><<<148, 128, smem>>>(d, 10);
  k<N, kTS, kAlt><<<148, 128, smem>>>(d, iters);
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double clk = (double)h[0] / (iters * 4.0);
  printf("%-40s N=%3d  %.1f clk per MMA  = %.0f FLOP/clk/SM  (%s)\n", name, N, clk, 2.0 * 128 * N * 16 / clk, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main() {
  run<192, false, 1>("SS, one accumulator");
  run<192, false, 2>("SS, two accumulators alternating");
  run<256, false, 1>("SS, one accumulator");
  run<128, false, 1>("SS, one accumulator");
  run<64, false, 1>("SS, one accumulator");
  run<192, true, 1>("TS (A in TMEM), one accumulator");
  run<128, true, 1>("TS (A in TMEM), one accumulator");
  run<208, true, 1>("TS (A in TMEM), one accumulator");
  run<112, true, 1>("TS (A in TMEM), one accumulator");
  return 0;
}
