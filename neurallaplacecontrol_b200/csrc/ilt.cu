// Generic Fourier-series inverse Laplace transform (the `fourier` ILT of torchlaplace as called from
// w_nl.py:137-144; PARITY UNPINNED, CPU statement in oracle/ilt.py:fourier_line_integrate):
//   x(t) = exp(gamma t)/T [ Re F_0 / 2 + sum_{k>=1} Re(F_k e^{i k pi t/T}) ],  T = 2(t+eps), gamma = alpha - ln(tol)/T
// over rows of S complex64 coefficients.  HBM-streaming kernel: 8*S + 8 algorithmic bytes per output.
//
// One persistent CTA per SM.  Rows are consumed in stages of R consecutive rows - one contiguous R*8*S-byte span - that
// an elected producer thread fetches with cp.async.bulk (TMA 1-D bulk copy) into a 3-deep shared-memory ring guarded by
// full/empty mbarriers, so ~200 KB per SM are in flight.  512 consumer threads: LPR = 512/R lanes share a row (k strided
// by LPR, then a shuffle reduction in a fixed order), which keeps all 16 warps busy for every S.
// Because T = 2(t+eps), pi t/T = pi/2 - delta with delta = pi eps / T tiny: e^{i k pi t/T} = i^k e^{-i k delta}; with
// LPR a multiple of 4 the exact factor i^k is a per-lane constant and e^{-i k delta} a two-term series (sincosf when
// k*delta is not small), so no phase error accumulates over k.
#include "common.cuh"
#include "tc_umma.cuh"

namespace nlc {

using namespace umma;

constexpr float kIltAlpha = 1.0e-3f, kIltEps = 1.0e-6f;
constexpr float kIltLnTol = -4.605170185988091f;  // ln(1e-2)
constexpr int kIltThreads = 512, kIltStages = 3, kIltTab = 512;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// partial sum over k = k0, k0+kstep, ... of Re(F_k i^k e^{-i k delta}); q = k0 & 3 is constant when kstep % 4 == 0.
// kSmall: k*delta <= 0.03 for every k, e^{-i a} = (1 - a^2/2) - i (a - a^3/6) with truncation < 4e-8.
template <int Q, bool kSmall, typename Ptr>
__device__ __forceinline__ float ilt_partial_q(Ptr coef, int S, int k0, int kstep, float delta) {
  float acc = 0.0f;
  float a = (float)k0 * delta;
  const float da = (float)kstep * delta;
#pragma unroll 4
  for (int k = k0; k < S; k += kstep, a += da) {
    const float2 f = coef[k];
    float c, s;
    if (kSmall) {
      c = fmaf(-0.5f * a, a, 1.0f);
      s = a * fmaf(a * a, -1.6666666667e-1f, 1.0f);
    } else {
      sincosf((float)k * delta, &s, &c);
    }
    // g = F_k i^k ; term = Re(g (c - i s)) = g.re c + g.im s
    if (Q == 0) acc += fmaf(f.x, c, f.y * s);
    if (Q == 1) acc += fmaf(f.x, s, -f.y * c);
    if (Q == 2) acc -= fmaf(f.x, c, f.y * s);
    if (Q == 3) acc += fmaf(f.y, c, -f.x * s);
  }
  return acc;
}
template <typename Ptr>
__device__ __forceinline__ float ilt_partial(Ptr coef, int S, int k0, int kstep, float delta, bool small) {
  if (small) {
    switch (k0 & 3) {
      case 0: return ilt_partial_q<0, true>(coef, S, k0, kstep, delta);
      case 1: return ilt_partial_q<1, true>(coef, S, k0, kstep, delta);
      case 2: return ilt_partial_q<2, true>(coef, S, k0, kstep, delta);
      default: return ilt_partial_q<3, true>(coef, S, k0, kstep, delta);
    }
  }
  switch (k0 & 3) {
    case 0: return ilt_partial_q<0, false>(coef, S, k0, kstep, delta);
    case 1: return ilt_partial_q<1, false>(coef, S, k0, kstep, delta);
    case 2: return ilt_partial_q<2, false>(coef, S, k0, kstep, delta);
    default: return ilt_partial_q<3, false>(coef, S, k0, kstep, delta);
  }
}

__global__ void __launch_bounds__(kIltThreads, 1) ilt_fourier_kernel(const float2* __restrict__ F, const float* __restrict__ tv,
                                                                    int t_per_row, long long n_rows, int n_t, int S, int R,
                                                                    float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);  // [kIltStages]
  uint64_t* empty = full + kIltStages;                     // [kIltStages]
  float* tab_delta = reinterpret_cast<float*>(smem_raw + 128);  // [kIltTab] per-time-index constants (shared time grid)
  float* tab_scale = tab_delta + kIltTab;
  unsigned char* stages = smem_raw + 128 + 2 * kIltTab * sizeof(float);
  const bool use_tab = !t_per_row && n_t <= kIltTab;
  if (use_tab)
    for (int j = threadIdx.x; j < n_t; j += kIltThreads) {
      const float t = tv[j], T = 2.0f * (t + kIltEps);
      tab_delta[j] = 3.14159265358979f * kIltEps / T;
      tab_scale[j] = expf((kIltAlpha - kIltLnTol / T) * t) / T;
    }
  const uint32_t stage_bytes = (uint32_t)R * 8u * (uint32_t)S;
  const int tid = threadIdx.x, lane = tid & 31;
  const int LPR = kIltThreads / R;           // lanes per row: 4, 8, 16 or 32
  const int sub = tid % LPR, rloc = tid / LPR;  // lane within the row group, row within the stage
  if (tid == 0) {
    for (int i = 0; i < kIltStages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, kIltThreads / 32); }
    mbar_fence_init();
  }
  __syncthreads();

  const long long n_full = n_rows / R;  // full stages
  const long long first = blockIdx.x, step = gridDim.x;
  const long long my_count = first < n_full ? (n_full - first + step - 1) / step : 0;

  // producer prologue: fill the ring
  if (tid == 0) {
    for (int i = 0; i < kIltStages && i < my_count; ++i) {
      mbar_expect_tx(full + i, stage_bytes);
      bulk_g2s(stages + (size_t)i * stage_bytes, reinterpret_cast<const unsigned char*>(F) + (size_t)(first + i * step) * stage_bytes,
               stage_bytes, full + i);
    }
  }
  for (long long it = 0; it < my_count; ++it) {
    const int st = (int)(it % kIltStages);
    const uint32_t par = (uint32_t)((it / kIltStages) & 1);
    const long long row = (first + it * step) * R + rloc;
    float t = 0.f, T = 1.f, delta, scale = 0.f;
    if (use_tab) {
      const int j = (int)(row % n_t);
      delta = tab_delta[j]; scale = tab_scale[j];
    } else {
      t = t_per_row ? __ldg(tv + row) : __ldg(tv + (int)(row % n_t));
      T = 2.0f * (t + kIltEps);
      delta = 3.14159265358979f * kIltEps / T;  // pi/2 - pi t/T
    }
    const bool small = (float)S * delta < 0.03f;
    mbar_wait(full + st, par);
    const float2* coef = reinterpret_cast<const float2*>(stages + (size_t)st * stage_bytes) + (size_t)rloc * S;
    float acc = ilt_partial(coef, S, sub, LPR, delta, small);
    if (sub == 0) acc -= 0.5f * coef[0].x;  // the k = 0 term enters with weight 1/2
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + st);  // this warp is done reading the stage
    for (int o = LPR >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (sub == 0) {
      if (!use_tab) scale = expf((kIltAlpha - kIltLnTol / T) * t) / T;
      out[row] = scale * acc;
    }
    // producer: refill this stage for iteration it + kIltStages once every warp has released it
    if (tid == 0 && it + kIltStages < my_count) {
      mbar_wait(empty + st, par);
      fence_proxy_async_smem();
      mbar_expect_tx(full + st, stage_bytes);
      bulk_g2s(stages + (size_t)st * stage_bytes,
               reinterpret_cast<const unsigned char*>(F) + (size_t)(first + (it + kIltStages) * step) * stage_bytes, stage_bytes, full + st);
    }
  }
  // tail rows (< R), straight from global memory, by the last CTA
  const long long tail0 = n_full * R;
  if (blockIdx.x == gridDim.x - 1 && tail0 + rloc < n_rows) {
    const long long row = tail0 + rloc;
    const float t = t_per_row ? tv[row] : tv[(int)(row % n_t)];
    const float T = 2.0f * (t + kIltEps);
    const float delta = 3.14159265358979f * kIltEps / T;
    const bool small = (float)S * delta < 0.03f;
    const float2* coef = F + (size_t)row * S;
    float acc = ilt_partial(coef, S, sub, LPR, delta, small);
    if (sub == 0) acc -= 0.5f * coef[0].x;
    // the shuffle below needs whole row groups: rows past n_rows simply do not exist in this warp's mask
    const unsigned mask = __activemask();
    for (int o = LPR >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(mask, acc, o);
    if (sub == 0) out[row] = expf((kIltAlpha - kIltLnTol / T) * t) / T * acc;
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Warp-private form (the default whenever >= 4 stage slots of 32 rows fit shared memory, S <= ~220).
//
// Every warp owns a private ring of NST stages of 32 consecutive rows and is its own TMA producer, so there is no
// cross-warp synchronisation at all: W * (NST - 1) * 32 * 8 * S bytes per SM stay in flight.  Lane == row: a lane walks
// its row with conflict-free shared-memory loads (odd S: 8-byte loads at the natural stride 2S words, one bulk copy per
// stage; even S: rows are copied one by one into a stride padded to 4 (mod 32) words and read with 16-byte loads), the
// k loop is unrolled by four so the factor i^k is compile-time, and four independent accumulators are combined in a
// fixed order.  32 consecutive outputs leave as one coalesced 128-byte store.
// ---------------------------------------------------------------------------------------------------------------
struct IltRowConst {
  float delta, scale;
};
__device__ __forceinline__ IltRowConst ilt_row_const(float t) {
  const float T = 2.0f * (t + kIltEps);
  IltRowConst r;
  r.delta = 3.14159265358979f * kIltEps / T;  // pi/2 - pi t/T
  r.scale = expf((kIltAlpha - kIltLnTol / T) * t) / T;
  return r;
}

// one term: acc_q += Re(F_k i^k e^{-i k delta}), q = k & 3 known at compile time
template <int Q, bool kSmall>
__device__ __forceinline__ void ilt_term(float x, float y, float kf, float delta, float& acc) {
  float c, s;
  if (kSmall) {
    const float a = kf * delta, a2 = a * a;
    c = fmaf(a2, -0.5f, 1.0f);
    s = a * fmaf(a2, -1.6666666667e-1f, 1.0f);
  } else {
    sincosf(kf * delta, &s, &c);
  }
  if (Q == 0) acc = fmaf(y, s, fmaf(x, c, acc));
  if (Q == 1) acc = fmaf(-y, c, fmaf(x, s, acc));
  if (Q == 2) acc = fmaf(-y, s, fmaf(-x, c, acc));
  if (Q == 3) acc = fmaf(-x, s, fmaf(y, c, acc));
}

// series of one row held in shared (or global) memory; VEC = 2 requires S even and a 16-byte aligned row
template <int VEC, bool kSmall>
__device__ __forceinline__ float ilt_row_sum(const float2* __restrict__ row, int S, float delta) {
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  int k = 0;
  if (VEC == 2) {
    const float4* row4 = reinterpret_cast<const float4*>(row);
#pragma unroll 2
    for (; k + 4 <= S; k += 4) {
      const float4 u = row4[k >> 1], v = row4[(k >> 1) + 1];
      const float kf = (float)k;
      ilt_term<0, kSmall>(u.x, u.y, kf, delta, acc0);
      ilt_term<1, kSmall>(u.z, u.w, kf + 1.0f, delta, acc1);
      ilt_term<2, kSmall>(v.x, v.y, kf + 2.0f, delta, acc2);
      ilt_term<3, kSmall>(v.z, v.w, kf + 3.0f, delta, acc3);
    }
  } else {
#pragma unroll 2
    for (; k + 4 <= S; k += 4) {
      const float2 f0 = row[k], f1 = row[k + 1], f2 = row[k + 2], f3 = row[k + 3];
      const float kf = (float)k;
      ilt_term<0, kSmall>(f0.x, f0.y, kf, delta, acc0);
      ilt_term<1, kSmall>(f1.x, f1.y, kf + 1.0f, delta, acc1);
      ilt_term<2, kSmall>(f2.x, f2.y, kf + 2.0f, delta, acc2);
      ilt_term<3, kSmall>(f3.x, f3.y, kf + 3.0f, delta, acc3);
    }
  }
  if (k < S) { const float2 f = row[k]; ilt_term<0, kSmall>(f.x, f.y, (float)k, delta, acc0); }
  if (k + 1 < S) { const float2 f = row[k + 1]; ilt_term<1, kSmall>(f.x, f.y, (float)(k + 1), delta, acc1); }
  if (k + 2 < S) { const float2 f = row[k + 2]; ilt_term<2, kSmall>(f.x, f.y, (float)(k + 2), delta, acc2); }
  return ((acc0 + acc1) + (acc2 + acc3)) - 0.5f * row[0].x;  // the k = 0 term enters with weight 1/2
}

template <int VEC>
__global__ void __launch_bounds__(512, 1) ilt_rows_kernel(const float2* __restrict__ F, const float* __restrict__ tv, int t_per_row,
                                                          long long n_rows, int n_t, int S, int stride_bytes, int NST,
                                                          float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full_all = reinterpret_cast<uint64_t*>(smem_raw);      // [W][NST] (<= 64 barriers)
  float* tab_delta = reinterpret_cast<float*>(smem_raw + 512);      // [kIltTab] per-time-index constants (shared grid)
  float* tab_scale = tab_delta + kIltTab;
  unsigned char* stages_all = smem_raw + 512 + 2 * kIltTab * sizeof(float);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, W = blockDim.x >> 5;
  const bool use_tab = !t_per_row && n_t <= kIltTab;
  if (use_tab)
    for (int j = tid; j < n_t; j += blockDim.x) {
      const IltRowConst rc = ilt_row_const(tv[j]);
      tab_delta[j] = rc.delta;
      tab_scale[j] = rc.scale;
    }
  uint64_t* full = full_all + warp * NST;
  if (lane == 0) {
    for (int i = 0; i < NST; ++i) mbar_init(full + i, 1);
    mbar_fence_init();
  }
  __syncthreads();

  const uint32_t row_bytes = 8u * (uint32_t)S;
  const uint32_t stage_bytes = 32u * (uint32_t)stride_bytes;
  unsigned char* stages = stages_all + (size_t)warp * NST * stage_bytes;
  const long long n_blocks = n_rows >> 5;  // full 32-row blocks
  const long long gw = (long long)blockIdx.x * W + warp, GW = (long long)gridDim.x * W;
  const long long my_count = gw < n_blocks ? (n_blocks - gw + GW - 1) / GW : 0;
  const unsigned char* Fb = reinterpret_cast<const unsigned char*>(F);

  auto issue = [&](long long i) {  // block gw + i*GW -> stage i % NST
    const int st = (int)(i % NST);
    const long long row0 = (gw + i * GW) << 5;
    unsigned char* dst = stages + (size_t)st * stage_bytes;
    if (VEC == 1) {
      if (lane == 0) {
        mbar_expect_tx(full + st, 32u * row_bytes);
        bulk_g2s(dst, Fb + (size_t)row0 * row_bytes, 32u * row_bytes, full + st);
      }
    } else {
      if (lane == 0) mbar_expect_tx(full + st, 32u * row_bytes);
      __syncwarp();
      bulk_g2s(dst + (size_t)lane * stride_bytes, Fb + (size_t)(row0 + lane) * row_bytes, row_bytes, full + st);
    }
  };
  for (int i = 0; i < NST && i < my_count; ++i) issue(i);

  int jb = (int)(((gw << 5)) % n_t);            // time index of the block's first row
  const int jstep = (int)((GW << 5) % n_t);
  for (long long it = 0; it < my_count; ++it) {
    const int st = (int)(it % NST);
    const uint32_t par = (uint32_t)((it / NST) & 1);
    const long long row = ((gw + it * GW) << 5) + lane;
    IltRowConst rc;
    if (use_tab) {
      const int j = (jb + lane) % n_t;
      rc.delta = tab_delta[j]; rc.scale = tab_scale[j];
    } else {
      rc = ilt_row_const(t_per_row ? __ldg(tv + row) : __ldg(tv + (jb + lane) % n_t));
    }
    jb += jstep; if (jb >= n_t) jb -= n_t;
    const bool small = __all_sync(0xffffffffu, (float)S * rc.delta < 0.03f);
    mbar_wait(full + st, par);
    const float2* rowp = reinterpret_cast<const float2*>(stages + (size_t)st * stage_bytes + (size_t)lane * stride_bytes);
    const float acc = small ? ilt_row_sum<VEC, true>(rowp, S, rc.delta) : ilt_row_sum<VEC, false>(rowp, S, rc.delta);
    out[row] = rc.scale * acc;
    __syncwarp();
    if (it + NST < my_count) {
      fence_proxy_async_smem();
      issue(it + NST);
    }
  }
  // tail rows (< 32) straight from global memory, by the last warp of the last CTA
  const long long tail0 = n_blocks << 5;
  if (blockIdx.x == gridDim.x - 1 && warp == W - 1 && tail0 + lane < n_rows) {
    const long long row = tail0 + lane;
    const IltRowConst rc = ilt_row_const(t_per_row ? tv[row] : tv[(int)(row % n_t)]);
    const float2* rowp = F + (size_t)row * S;
    const float acc = ((float)S * rc.delta < 0.03f) ? ilt_row_sum<1, true>(rowp, S, rc.delta) : ilt_row_sum<1, false>(rowp, S, rc.delta);
    out[row] = rc.scale * acc;
  }
}

}  // namespace nlc

using namespace nlc;

extern "C" int nlc_ilt_fourier(const float* F_dev, const float* t_dev, int t_per_row, int64_t N, int n_t, int S,
                               float* out_dev, void* stream) {
  NLC_REQUIRE(F_dev && t_dev && out_dev, NLC_ERR_ARG, "nlc_ilt_fourier: null pointer");
  NLC_REQUIRE(N >= 1 && n_t >= 1 && S >= 1, NLC_ERR_ARG, "nlc_ilt_fourier: N, n_t, S must be positive");
  NLC_REQUIRE(S <= 512, NLC_ERR_SHAPE, "nlc_ilt_fourier: S = %d exceeds 512", S);
  NLC_REQUIRE((reinterpret_cast<uintptr_t>(F_dev) & 15) == 0, NLC_ERR_ARG, "F_dev must be 16-byte aligned");
  const long long n_rows = (long long)N * n_t;
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  // warp-private rings: stride of a row in shared memory (even S: padded to 4 (mod 32) words for 16-byte loads)
  {
    const int vec = (S % 2 == 0) ? 2 : 1;
    int stride = 8 * S;
    if (vec == 2) stride += 16 * ((((1 - S / 2) % 8) + 8) % 8);  // m sixteen-byte pads with S/2 + m = 1 (mod 8)
    const size_t stage = (size_t)32 * stride;
    const size_t fixed = 512 + 2 * kIltTab * sizeof(float);
    const int slots = (int)((227 * 1024 - fixed) / stage);
    if (slots >= 4 && n_rows >= 32) {
      const int nst = slots >= 6 ? 3 : 2;
      int W = slots / nst;
      if (W > 8) W = 8;
      const size_t smem = fixed + (size_t)W * nst * stage;
      void (*kern)(const float2*, const float*, int, long long, int, int, int, int, float*) = vec == 2 ? ilt_rows_kernel<2> : ilt_rows_kernel<1>;
      NLC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
      const long long n_blocks = n_rows / 32;
      long long grid = (n_blocks + W - 1) / W;
      if (grid > 148) grid = 148;
      kern<<<(int)grid, W * 32, smem, cs>>>(reinterpret_cast<const float2*>(F_dev), t_dev, t_per_row, n_rows, n_t, S, stride, nst, out_dev);
      NLC_LAUNCH_OK("ilt_rows_kernel");
      return NLC_OK;
    }
  }
  // CTA-wide ring for long rows.  rows per stage: the largest of 128/64/32/16 whose 3-stage ring fits
  int R = 128;
  while (R > 16 && (size_t)kIltStages * R * 8 * S > 212 * 1024) R >>= 1;
  NLC_REQUIRE((size_t)kIltStages * R * 8 * S <= 212 * 1024, NLC_ERR_SHAPE, "nlc_ilt_fourier: stage does not fit shared memory");
  const size_t smem = 128 + 2 * kIltTab * sizeof(float) + (size_t)kIltStages * R * 8 * S;
  NLC_CUDA_OK(cudaFuncSetAttribute(ilt_fourier_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
  const long long n_full = n_rows / R;
  long long grid = 148;
  if (n_full < grid) grid = n_full > 0 ? n_full : 1;
  ilt_fourier_kernel<<<(int)grid, kIltThreads, smem, cs>>>(
      reinterpret_cast<const float2*>(F_dev), t_dev, t_per_row, n_rows, n_t, S, R, out_dev);
  NLC_LAUNCH_OK("ilt_fourier_kernel");
  return NLC_OK;
}
