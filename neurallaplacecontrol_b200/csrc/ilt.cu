// Generic Fourier-series inverse Laplace transform (the `fourier` ILT of torchlaplace as called from
// w_nl.py:137-144; PARITY UNPINNED, CPU statement in oracle/ilt.py:fourier_line_integrate):
//   x(t) = exp(gamma t)/T [ Re F_0 / 2 + sum_{k>=1} Re(F_k e^{i k pi t/T}) ],  T = 2(t+eps), gamma = alpha - ln(tol)/T
// over rows of S complex64 coefficients.  HBM-streaming kernel: 8*S + 8 algorithmic bytes per output.
//
// Each warp owns 32 consecutive rows at a time - a contiguous 256*S-byte span - fetched with ONE
// cp.async.bulk (TMA 1-D bulk copy, completion on an mbarrier) into a warp-private, double-buffered
// shared-memory stage; lane l then sums row l in fixed k order (deterministic), reading 8-byte
// coefficients at a stride of 2S words, which is bank-conflict-free for odd S (an index rotation makes
// it so for even S).  Because T = 2(t+eps), pi t/T = pi/2 - delta with delta = pi eps / T tiny:
// e^{i k pi t/T} = i^k e^{-i k delta}; i^k is exact and e^{-i k delta} is a short Taylor series
// (sincosf when k*delta is not small), so no phase error accumulates over k.
#include "common.cuh"

namespace nlc {

constexpr float kIltAlpha = 1.0e-3f, kIltEps = 1.0e-6f;
constexpr float kIltLnTol = -4.605170185988091f;  // ln(1e-2)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// one row: coef points at S complex values (smem or global), stride in float2 units between k
template <typename Ptr>
__device__ __forceinline__ float ilt_row(Ptr coef, int S, int rot, float t) {
  const float T = 2.0f * (t + kIltEps);
  const float gamma = kIltAlpha - kIltLnTol / T;
  const float delta = 3.14159265358979f * kIltEps / T;  // pi/2 - pi t/T = pi eps / T
  const bool small = (float)S * delta < 0.03f;
  float acc = 0.0f;
  for (int kk = 0; kk < S; ++kk) {
    int k = kk + rot;
    if (k >= S) k -= S;
    const float2 f = coef[k];
    const float ang = (float)k * delta;
    float c, s;
    if (small) {
      const float a2 = ang * ang;
      c = fmaf(a2, fmaf(a2, 4.1666666667e-2f, -0.5f), 1.0f);
      s = ang * fmaf(a2, fmaf(a2, 8.3333333333e-3f, -1.6666666667e-1f), 1.0f);
    } else {
      sincosf(ang, &s, &c);
    }
    // g = F_k * i^k
    const int q = k & 3;
    const float gr = q == 0 ? f.x : (q == 1 ? -f.y : (q == 2 ? -f.x : f.y));
    const float gi = q == 0 ? f.y : (q == 1 ? f.x : (q == 2 ? -f.y : -f.x));
    float term = fmaf(gr, c, gi * s);  // Re(g (c - i s))
    if (k == 0) term *= 0.5f;
    acc += term;
  }
  return expf(gamma * t) / T * acc;
}

__global__ void __launch_bounds__(256) ilt_fourier_kernel(const float2* __restrict__ F, const float* __restrict__ tv,
                                                          int t_per_row, long long n_rows, int n_t, int S,
                                                          float* __restrict__ out, int warps_per_cta) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t stage_bytes = 256u * (uint32_t)S;  // 32 rows
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);  // [warps][2]
  unsigned char* stages = smem_raw + 128 + (size_t)warp * 2 * stage_bytes;
  uint64_t* bar = bars + warp * 2;
  if (lane == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();

  const long long n_chunks = n_rows >> 5;  // full 32-row chunks
  const long long gw = (long long)blockIdx.x * warps_per_cta + warp;
  const long long total_warps = (long long)gridDim.x * warps_per_cta;
  // even S: rotate each lane's starting k so that lanes hit distinct banks
  const int q = ((S & 1) == 0) ? ((17 - (S & 15)) & 15) : 0;
  const int rot = (lane * q) % S;

  long long c = gw;
  uint32_t phase0 = 0, phase1 = 0;
  if (c < n_chunks && lane == 0) {
    mbar_expect_tx(bar, stage_bytes);
    bulk_g2s(stages, reinterpret_cast<const unsigned char*>(F) + (size_t)c * stage_bytes, stage_bytes, bar);
  }
  int st = 0;
  for (; c < n_chunks; c += total_warps, st ^= 1) {
    const long long nxt = c + total_warps;
    if (nxt < n_chunks && lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(bar + (st ^ 1), stage_bytes);
      bulk_g2s(stages + (size_t)(st ^ 1) * stage_bytes, reinterpret_cast<const unsigned char*>(F) + (size_t)nxt * stage_bytes,
               stage_bytes, bar + (st ^ 1));
    }
    const long long row = (c << 5) + lane;
    const float t = t_per_row ? __ldg(tv + row) : __ldg(tv + (int)(row % n_t));
    mbar_wait(bar + st, st ? phase1 : phase0);
    if (st) phase1 ^= 1; else phase0 ^= 1;
    const float2* coef = reinterpret_cast<const float2*>(stages + (size_t)st * stage_bytes) + (size_t)lane * S;
    out[row] = ilt_row(coef, S, rot, t);
    __syncwarp();  // every lane is done with this stage before it is refilled
  }
  // tail rows (< 32), straight from global memory, by the warp that would own the next chunk
  const long long tail0 = n_chunks << 5;
  if (tail0 < n_rows && gw == (n_chunks % total_warps)) {
    const long long row = tail0 + lane;
    if (row < n_rows) {
      const float t = t_per_row ? tv[row] : tv[(int)(row % n_t)];
      out[row] = ilt_row(F + (size_t)row * S, S, 0, t);
    }
  }
}

}  // namespace nlc

using namespace nlc;

extern "C" int nlc_ilt_fourier(const float* F_dev, const float* t_dev, int t_per_row, int64_t N, int n_t, int S,
                               float* out_dev, void* stream) {
  NLC_REQUIRE(F_dev && t_dev && out_dev, NLC_ERR_ARG, "nlc_ilt_fourier: null pointer");
  NLC_REQUIRE(N >= 1 && n_t >= 1 && S >= 1, NLC_ERR_ARG, "nlc_ilt_fourier: N, n_t, S must be positive");
  NLC_REQUIRE(S <= 448, NLC_ERR_SHAPE, "nlc_ilt_fourier: S = %d exceeds 448", S);
  NLC_REQUIRE((reinterpret_cast<uintptr_t>(F_dev) & 15) == 0, NLC_ERR_ARG, "F_dev must be 16-byte aligned");
  const size_t stage_bytes = 256 * (size_t)S;
  int warps = (int)((200 * 1024) / (2 * stage_bytes));
  if (warps > 8) warps = 8;
  NLC_REQUIRE(warps >= 1, NLC_ERR_SHAPE, "nlc_ilt_fourier: stage does not fit shared memory");
  const size_t smem = 128 + (size_t)warps * 2 * stage_bytes;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    NLC_CUDA_OK(cudaFuncSetAttribute(ilt_fourier_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
    attr_smem = 227 * 1024;
  }
  const long long n_rows = (long long)N * n_t;
  const long long n_chunks = n_rows >> 5;
  int ctas_per_sm = (int)((227 * 1024) / (smem + 1024));
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  if (ctas_per_sm > 4) ctas_per_sm = 4;
  long long grid = 148LL * ctas_per_sm;
  const long long need = (n_chunks + warps - 1) / warps;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  ilt_fourier_kernel<<<(int)grid, warps * 32, smem, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2*>(F_dev), t_dev, t_per_row, n_rows, n_t, S, out_dev, warps);
  NLC_LAUNCH_OK("ilt_fourier_kernel");
  return NLC_OK;
}
