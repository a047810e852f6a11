// Self-tests of the tcgen05 operand layouts / descriptors / TMEM addressing the tensor-core kernels rely on (no reference
// counterpart): a single 128 x N x K product with operands split hi+lo on the device exactly as the kernels do, with the
// A operand in shared memory (encode_tc2.cu's path) or in tensor memory (rollout_tc2.cu's path).
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_pack.cuh"
#include "tc_umma.cuh"

namespace nlc {

using namespace umma;

constexpr int kTcRows = 128;
constexpr int kTcHg = 64;
constexpr uint32_t kOpBytes = kTcRows * kTcHg * 2;  // one fp16 A-operand image (16 KB)
constexpr uint32_t kLbo = 128, kSbo = (kTcHg / 8) * 128;  // K-major no-swizzle, K = 64

// 4 (x3) MMAs: D[128 x N] (+)= A[128 x 64] * B[N x 64]^T
template <bool kSplit3>
__device__ __forceinline__ void issue_gemm(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                           int N, bool accumulate) {
  const uint32_t idesc = idesc_f16_f32(kTcRows, N);
#pragma unroll
  for (int ks = 0; ks < kTcHg / 16; ++ks) {
    const uint32_t off = ks * 2 * kLbo;  // 16 K-elements = two core matrices
    mma_f16_ss(d_tmem, smem_desc(a_hi + off, kLbo, kSbo), smem_desc(b_hi + off, kLbo, kSbo), idesc, (accumulate || ks > 0) ? 1u : 0u);
    if (kSplit3) {
      mma_f16_ss(d_tmem, smem_desc(a_lo + off, kLbo, kSbo), smem_desc(b_hi + off, kLbo, kSbo), idesc, 1u);
      mma_f16_ss(d_tmem, smem_desc(a_hi + off, kLbo, kSbo), smem_desc(b_lo + off, kLbo, kSbo), idesc, 1u);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Self-test of the operand layout / descriptors / TMEM addressing: D[128][N] = A[128][64] B[n_off : n_off+N][64]^T
// with A, B given in fp32 and split on the device exactly as the encoder does.
// ---------------------------------------------------------------------------------------------------------------
template <bool kSplit3>
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                               int n_rows_b, int n_off, int N, float* __restrict__ D) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* a_img[2] = {smem_raw, smem_raw + kOpBytes};
  unsigned char* b_img[2] = {smem_raw + 2 * kOpBytes, smem_raw + 2 * kOpBytes + 256 * kTcHg * 2};
  __shared__ alignas(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < kTcRows * kTcHg; i += 128) {
    const int r = i / kTcHg, k = i - r * kTcHg;
    const float v = A[i];
    const __half h = __float2half_rn(v);
    reinterpret_cast<__half*>(a_img[0])[tc_core_offset(r, k, kTcHg)] = h;
    reinterpret_cast<__half*>(a_img[1])[tc_core_offset(r, k, kTcHg)] = __float2half_rn(v - __half2float(h));
  }
  for (int i = tid; i < n_rows_b * kTcHg; i += 128) {
    const int r = i / kTcHg, k = i - r * kTcHg;
    const float v = Bm[i];
    const __half h = __float2half_rn(v);
    reinterpret_cast<__half*>(b_img[0])[tc_core_offset(r, k, kTcHg)] = h;
    reinterpret_cast<__half*>(b_img[1])[tc_core_offset(r, k, kTcHg)] = __float2half_rn(v - __half2float(h));
  }
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 256);
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    const uint32_t boff = (uint32_t)(n_off / 8) * kSbo;
    issue_gemm<kSplit3>(tmem, smem_u32(a_img[0]), smem_u32(a_img[1]), smem_u32(b_img[0]) + boff, smem_u32(b_img[1]) + boff, N, false);
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  const uint32_t tlane = tmem + ((uint32_t)(32 * warp) << 16);
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tlane + c0, v);
    tmem_ld_wait();
    for (int i = 0; i < 16 && c0 + i < N; ++i) D[(size_t)(32 * warp + lane) * N + c0 + i] = v[i];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// Same product with the A operand staged in TENSOR MEMORY (tcgen05.st by the owning threads, tcgen05.mma reading
// [a_tmem]): the operand path of the fused rollout kernel.  K = 128 here (two 64-wide halves of A are given).
template <bool kSplit3>
__global__ void __launch_bounds__(128, 1) umma_selftest_ts_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                                  int n_rows_b, int N, int Kdim, float* __restrict__ D) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* b_img[2] = {smem_raw, smem_raw + 256 * 128 * 2};
  __shared__ alignas(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < n_rows_b * Kdim; i += 128) {
    const int r = i / Kdim, k = i - r * Kdim;
    const float v = Bm[i];
    const __half h = __float2half_rn(v);
    reinterpret_cast<__half*>(b_img[0])[tc_core_offset(r, k, Kdim)] = h;
    reinterpret_cast<__half*>(b_img[1])[tc_core_offset(r, k, Kdim)] = __float2half_rn(v - __half2float(h));
  }
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base;
  const uint32_t tlane = tmem + ((uint32_t)(32 * warp) << 16);
  // A operand: columns [256, 256+Kdim/2) hi, [384, 384+Kdim/2) lo ; accumulator at columns [0, N)
  const uint32_t colAhi = 256, colAlo = 384;
  const int row = 32 * warp + lane;
  for (int c0 = 0; c0 < Kdim / 2; c0 += 16) {
    uint32_t ph[16], pl[16];
    for (int i = 0; i < 16; ++i) {
      const float x0 = A[(size_t)row * Kdim + 2 * (c0 + i)], x1 = A[(size_t)row * Kdim + 2 * (c0 + i) + 1];
      const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
      const __half2 hh = __halves2half2(h0, h1);
      const __half2 ll = __halves2half2(__float2half_rn(x0 - __half2float(h0)), __float2half_rn(x1 - __half2float(h1)));
      ph[i] = *reinterpret_cast<const uint32_t*>(&hh);
      pl[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    tmem_st16(tlane + colAhi + c0, ph);
    tmem_st16(tlane + colAlo + c0, pl);
  }
  tmem_st_wait();
  fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    fence_after_sync();
    const uint32_t idesc = idesc_f16_f32(kTcRows, N);
    const uint32_t sbo = (uint32_t)(Kdim / 8) * 128;
    for (int ks = 0; ks < Kdim / 16; ++ks) {
      const uint32_t boff = ks * 2 * kLbo;
      mma_f16_ts(tmem, tmem + colAhi + 8 * ks, smem_desc(smem_u32(b_img[0]) + boff, kLbo, sbo), idesc, ks > 0 ? 1u : 0u);
      if (kSplit3) {
        mma_f16_ts(tmem, tmem + colAlo + 8 * ks, smem_desc(smem_u32(b_img[0]) + boff, kLbo, sbo), idesc, 1u);
        mma_f16_ts(tmem, tmem + colAhi + 8 * ks, smem_desc(smem_u32(b_img[1]) + boff, kLbo, sbo), idesc, 1u);
      }
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tlane + c0, v);
    tmem_ld_wait();
    for (int i = 0; i < 16 && c0 + i < N; ++i) D[(size_t)row * N + c0 + i] = v[i];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace nlc

using namespace nlc;

extern "C" int nlc_selftest_umma_gemm_ts(const float* A_dev, const float* B_dev, int n_rows_b, int N, int Kdim, int split3,
                                         float* D_dev, void* stream) {
  NLC_REQUIRE(A_dev && B_dev && D_dev, NLC_ERR_ARG, "nlc_selftest_umma_gemm_ts: null pointer");
  NLC_REQUIRE(n_rows_b % 8 == 0 && n_rows_b <= 256 && N % 16 == 0 && N >= 16 && N <= n_rows_b && (Kdim == 64 || Kdim == 128),
              NLC_ERR_SHAPE, "nlc_selftest_umma_gemm_ts: bad shape");
  const int smem = 2 * 256 * 128 * 2 + 128;
  NLC_CUDA_OK(cudaFuncSetAttribute(umma_selftest_ts_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  NLC_CUDA_OK(cudaFuncSetAttribute(umma_selftest_ts_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (split3) umma_selftest_ts_kernel<true><<<1, 128, smem, s>>>(A_dev, B_dev, n_rows_b, N, Kdim, D_dev);
  else umma_selftest_ts_kernel<false><<<1, 128, smem, s>>>(A_dev, B_dev, n_rows_b, N, Kdim, D_dev);
  NLC_LAUNCH_OK("umma_selftest_ts_kernel");
  return NLC_OK;
}

extern "C" int nlc_selftest_umma_gemm(const float* A_dev, const float* B_dev, int n_rows_b, int n_off, int N, int split3,
                                      float* D_dev, void* stream) {
  NLC_REQUIRE(A_dev && B_dev && D_dev, NLC_ERR_ARG, "nlc_selftest_umma_gemm: null pointer");
  NLC_REQUIRE(n_rows_b % 8 == 0 && n_rows_b <= 256 && n_off % 8 == 0 && N % 16 == 0 && N >= 16 && n_off + N <= n_rows_b,
              NLC_ERR_SHAPE, "nlc_selftest_umma_gemm: bad shape");
  const int smem = 2 * kOpBytes + 2 * 256 * kTcHg * 2 + 128;
  NLC_CUDA_OK(cudaFuncSetAttribute(umma_selftest_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  NLC_CUDA_OK(cudaFuncSetAttribute(umma_selftest_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (split3) umma_selftest_kernel<true><<<1, 128, smem, s>>>(A_dev, B_dev, n_rows_b, n_off, N, D_dev);
  else umma_selftest_kernel<false><<<1, 128, smem, s>>>(A_dev, B_dev, n_rows_b, n_off, N, D_dev);
  NLC_LAUNCH_OK("umma_selftest_kernel");
  return NLC_OK;
}
