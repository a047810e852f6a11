// Stages 2+3 for the Neural Laplace dynamics, representation MLP on the 5th-generation tensor cores
// (rollout.cu is the fp32 CUDA-core anchor).
//
// Same recurrence as rollout.cu: state <- state + ILT(MLP([s | obs_n | p_action])), cost += running_cost, whole horizon in
// one launch, state in registers.  The recurrence is a strict chain per sample (L1 -> L2 -> L3 -> ILT -> next step), so
// inside ONE tile the tensor pipe and the CUDA cores can only alternate; the first form measured 22.7 k clocks per step
// = MUFU time + issue time + MMA time, nothing overlapping.  Here every CTA runs TWO independent tiles ("groups" of 8
// warps = 128 samples each) that share the weight images in shared memory and otherwise never synchronise with each
// other: one group's MMAs and barrier round trips run under the other group's epilogues.  (Holding the two groups
// exactly half a step apart with named barriers was tried: 29.5 k instead of 31.3 k clocks per step in the trace, but
// slower end to end - tools/trace_rollout.py shows long transients with every phase stretched - so they run free.)
//
//   per step and group      A1 = [obs_n | p_action | 1] as a K = 16 operand (TMEM)          (one thread per sample)
//     M1  D = A1 W1^T       N = 128, K = 16   the first layer WITH its folded bias on the tensor cores
//     E1  tanh -> A         hidden activations re-written as the fp16 hi/lo A operand in TMEM (tcgen05.st)
//     M2  D = A W2^T        N = 128, K = 128
//     E2  + b2, tanh -> A
//     M3a D = A W3[:N3a]^T  the (theta, phi) pre-activations in two column halves so that A (128 columns) and D
//     E3a                   (128 columns) of BOTH groups fit the 512 TMEM columns
//     M3b / E3b             second half (skipped when 2 nx S <= 128)
//     E3: sphere -> complex map, Fourier weights, fixed-order sum over the terms; the two column groups of a sample
//     exchange partial sums through shared memory; residual add; running cost (overlapped with M3a)
//
//   -2 log2(e) is folded into all three layers on the host (model.cu): tanh(x) = 2 / (1 + 2^x') - 1 with x' = -2 log2e x,
//   the reciprocals are Newton iterations on the FMA pipe in packed fp32x2 (f32x2.cuh); per (channel, term) pair the
//   epilogue needs 2 ex2 + 1 cos + 1 rcp on the MUFU pipe instead of 7.
//   NLC_MATH_TC_SPLIT3: operands split hi+lo in fp16, D += A_hi B_hi + A_lo B_hi + A_hi B_lo (fp32-class);
//   NLC_MATH_TC_FP16  : single pass, tanh.approx (the stated looser bound).
//
//   Samples are dealt to the 2 x gridDim groups in contiguous ranges (multiples of 32 rows), each walked in tiles of up
//   to 128 rows: a plan that is not a whole number of waves ends on partially filled tiles instead of an idle wave.
//   Threads: group g = warps 8g .. 8g+7; warp w of a group owns TMEM lanes 32 (w & 3).. (its 32 samples) and column
//   half w >> 2.  After a group barrier its first warp issues the group's MMAs: the group waits for that product anyway,
//   so the issuing thread being held by the MMA queue costs nothing (unlike in the encoder).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "env_cost.cuh"
#include "f32x2.cuh"
#include "tc_umma.cuh"

namespace nlc {

using namespace umma;

namespace rt2 {

constexpr int kRows = 128, kH = 128;
constexpr uint32_t kLbo = 128, kSbo = (kH / 8) * 128;  // K = 128 operand images
constexpr uint32_t kSbo16 = (16 / 8) * 128;             // K = 16 (first layer)
constexpr uint32_t kColA = 0, kColD = 128, kGroupCols = 256, kTmemCols = 512;
constexpr int kGroupThreads = 256, kThreads = 2 * kGroupThreads;

struct Args {
  ModelDev m;
  nlc_rollout_opts o;
  const float* state0; int state_per_sample;
  const float* p;
  const float* hist;
  const float* pert_cost;
  int K, T, B, L, nu;
  float* cost_total;
  float* states;
  float* delta_out;
  long long* trace;  // measurement only: clock64 timeline of CTA 0, [step][warp][8 events]
};

struct SmemTail {  // after the weight images
  alignas(16) float b2[kH];
  alignas(16) float b3[256];
  float phase[kMaxS], weight[kMaxS];
  float smean[kMaxNx], sinv[kMaxNx];
  alignas(16) float exch[4][kRows * kMaxNx];     // [tile * column groups + column group] partial ILT sums
  alignas(8) uint64_t done[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float mufu_tanh(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_cos(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// tanh of two pre-activations that carry the folded scale x' = -2 log2(e) x
template <bool kAccurate, int kRcp>
__device__ __forceinline__ f2_t tanh2_scaled(f2_t xs) {
  float a, b;
  if (kAccurate) {
    upk2(xs, a, b);
    const f2_t d = add2(pk2(mufu_ex2(fminf(a, 120.0f)), mufu_ex2(fminf(b, 120.0f))), pk2(1.0f, 1.0f));
    return fma2(pk2(2.0f, 2.0f), rcp2<kRcp>(d), pk2(-1.0f, -1.0f));
  }
  upk2(mul2(xs, pk2(-0.34657359027997264f, -0.34657359027997264f)), a, b);
  return pk2(mufu_tanh(a), mufu_tanh(b));
}

// 16 fp32 values (8 packed pairs) -> 8 fp16x2 words (hi) and the fp16 residuals (lo)
template <bool kSplit3>
__device__ __forceinline__ void pack16(const f2_t (&v)[8], uint32_t (&ph)[8], uint32_t (&pl)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float x0, x1;
    upk2(v[i], x0, x1);
    const __half2 hh = __floats2half2_rn(x0, x1);
    ph[i] = *reinterpret_cast<const uint32_t*>(&hh);
    if (kSplit3) {
      const float2 back = __half22float2(hh);
      const __half2 ll = __floats2half2_rn(x0 - back.x, x1 - back.y);
      pl[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
  }
}
__device__ __forceinline__ void ldtm16p(uint32_t taddr, f2_t (&v)[8]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = pk2u(r[2 * i], r[2 * i + 1]);
}

// D[128 x N] = A[128 x Kdim] (TMEM, hi at a_tmem, lo at a_tmem + 64) * B[N x Kdim]^T (smem image, K-major)
template <bool kSplit3>
__device__ __forceinline__ void issue_gemm_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_hi, uint32_t b_lo, int N, int ksteps, uint32_t sbo) {
  const uint32_t idesc = idesc_f16_f32(kRows, N);
  for (int ks = 0; ks < ksteps; ++ks) {
    const uint32_t boff = ks * 2 * kLbo;
    mma_f16_ts(d_tmem, a_tmem + 8 * ks, smem_desc(b_hi + boff, kLbo, sbo), idesc, ks > 0 ? 1u : 0u);
    if (kSplit3) {
      mma_f16_ts(d_tmem, a_tmem + 64 + 8 * ks, smem_desc(b_hi + boff, kLbo, sbo), idesc, 1u);
      mma_f16_ts(d_tmem, a_tmem + 8 * ks, smem_desc(b_lo + boff, kLbo, sbo), idesc, 1u);
    }
  }
}

// Two (channel, term) pairs of the L3 epilogue.  v = (theta', phi') pre-activations (bias added, scale folded):
//   y = tanh(theta_pre) -> theta = pi y (w_nl.py:59);   radius = tan((pi/2) sigmoid(2 phi_pre)) (w_nl.py:60-62 + sphere map,
//   written through s = sigmoid(-2|phi_pre|) in (0, 1/2] so both tails keep relative accuracy, common.cuh:sphere_radius);
//   term = weight_k radius cos(pi y + k pi t/T).      MUFU: 2 ex2 + 1 cos + 1 rcp per pair; the two sigmoids' reciprocals
//   come from ONE Newton reciprocal of the product of their denominators.
template <bool kAccurate, int kRcp>
__device__ __forceinline__ void l3_two_pairs(f2_t th, f2_t ph, float phase0, float phase1, float w0, float w1, float& t0, float& t1) {
  float a0, a1, b0, b1;
  upk2(th, a0, a1);
  upk2(ph, b0, b1);
  const f2_t one = pk2(1.0f, 1.0f);
  f2_t y, sg;
  const f2_t e2 = pk2(mufu_ex2(-fabsf(b0)), mufu_ex2(-fabsf(b1)));  // exp(-2|phi_pre|) in (0, 1]
  if (kAccurate) {
    const f2_t d1 = add2(pk2(mufu_ex2(fminf(a0, 60.0f)), mufu_ex2(fminf(a1, 60.0f))), one);
    const f2_t d2 = add2(e2, one);
    if (kRcp == 0) {  // two independent MUFU reciprocals: shortest dependency chain
      y = fma2(pk2(2.0f, 2.0f), rcp2<0>(d1), pk2(-1.0f, -1.0f));
      sg = mul2(e2, rcp2<0>(d2));
    } else {          // one Newton reciprocal of the product on the FMA pipe: fewest MUFU operations
      const f2_t inv = rcp2<kRcp>(mul2(d1, d2));
      y = fma2(pk2(2.0f, 2.0f), mul2(d2, inv), pk2(-1.0f, -1.0f));
      sg = mul2(e2, mul2(d1, inv));
    }
  } else {
    float c0, c1;
    upk2(mul2(th, pk2(-0.34657359027997264f, -0.34657359027997264f)), c0, c1);
    y = pk2(mufu_tanh(c0), mufu_tanh(c1));
    float d0, d1_;
    upk2(add2(e2, one), d0, d1_);
    sg = mul2(e2, pk2(mufu_rcp(d0), mufu_rcp(d1_)));
  }
  const f2_t x = mul2(sg, pk2(1.57079632679489662f, 1.57079632679489662f));  // (0, pi/4]
  const f2_t x2 = mul2(x, x);
  f2_t sn = fma2(x2, pk2(2.7557319224e-6f, 2.7557319224e-6f), pk2(-1.9841269841e-4f, -1.9841269841e-4f));
  sn = fma2(sn, x2, pk2(8.3333333333e-3f, 8.3333333333e-3f));
  sn = fma2(sn, x2, pk2(-1.6666666667e-1f, -1.6666666667e-1f));
  sn = mul2(x, fma2(sn, x2, one));
  f2_t cs = fma2(x2, pk2(2.4801587302e-5f, 2.4801587302e-5f), pk2(-1.3888888889e-3f, -1.3888888889e-3f));
  cs = fma2(cs, x2, pk2(4.1666666667e-2f, 4.1666666667e-2f));
  cs = fma2(cs, x2, pk2(-0.5f, -0.5f));
  cs = fma2(cs, x2, one);
  float sn0, sn1, cs0, cs1;
  upk2(sn, sn0, sn1);
  upk2(cs, cs0, cs1);
  // phi_pre <= 0 (scaled value >= 0): radius = tan(x) = sn / cs, else cot(x) = cs / sn
  const bool neg0 = b0 >= 0.0f, neg1 = b1 >= 0.0f;
  const float rad0 = (neg0 ? sn0 : cs0) * mufu_rcp(neg0 ? cs0 : sn0);
  const float rad1 = (neg1 ? sn1 : cs1) * mufu_rcp(neg1 ? cs1 : sn1);
  // cos(pi y + a_k), a_k = pi k t/T reduced to (-pi, pi]: the argument stays inside (-2 pi, 2 pi), where cos.approx is
  // good to 5.2e-7 absolute without any further range reduction (tools/cos_err.cu; 3.4e-7 with an exact reduction)
  float y0, y1;
  upk2(y, y0, y1);
  t0 = (w0 * rad0) * mufu_cos(fmaf(3.14159265358979f, y0, phase0));
  t1 = (w1 * rad1) * mufu_cos(fmaf(3.14159265358979f, y1, phase1));
}

// L3 epilogue of one 16-column chunk (8 pairs); kChunk = chunk index in the full N3t column space, kCol0 = first chunk
// of the half that currently sits in the D region
template <int NX, int S, int kChunk, int kCol0, bool kAccurate, int kRcp>
__device__ __forceinline__ void l3_chunk(uint32_t tD, const float* __restrict__ b3, const float* __restrict__ phase,
                                         const float* __restrict__ weight, float (&delta)[NX]) {
  f2_t v[8];
  ldtm16p(tD + 16 * (kChunk - kCol0), v);
  tmem_ld_wait();
  f2_t bb[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = *reinterpret_cast<const float4*>(b3 + 16 * kChunk + 4 * i);
    bb[2 * i] = pk2(t.x, t.y);
    bb[2 * i + 1] = pk2(t.z, t.w);
  }
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    constexpr int dummy = 0; (void)dummy;
    const int p0 = 8 * kChunk + i, p1 = p0 + 1;
    if (p0 < NX * S) {
      // v[i] = (theta', phi') of pair p0, v[i+1] of pair p1: regroup into (theta0, theta1), (phi0, phi1)
      const f2_t s0 = add2(v[i], bb[i]), s1 = add2(v[i + 1], bb[i + 1]);
      float th0, ph0, th1, ph1;
      upk2(s0, th0, ph0);
      upk2(s1, th1, ph1);
      const int ch0 = p0 / S, k0 = p0 - ch0 * S;
      const int ch1 = (p1 < NX * S) ? p1 / S : ch0, k1 = (p1 < NX * S) ? p1 - ch1 * S : k0;
      float t0, t1;
      l3_two_pairs<kAccurate, kRcp>(pk2(th0, th1), pk2(ph0, ph1), phase[k0], phase[k1], weight[k0], weight[k1], t0, t1);
      delta[ch0] += t0;
      if (p1 < NX * S) delta[ch1] += t1;
    }
  }
}
// chunks kChunk, kChunk + kStride, ... < kEnd (the column groups of a sample take chunks round-robin)
template <int NX, int S, int kChunk, int kEnd, int kCol0, bool kAccurate, int kRcp, int kStride>
struct L3Loop {
  static __device__ __forceinline__ void run(uint32_t tD, const float* b3, const float* phase, const float* weight, float (&delta)[NX]) {
    if constexpr (kChunk < kEnd) {
      l3_chunk<NX, S, kChunk, kCol0, kAccurate, kRcp>(tD, b3, phase, weight, delta);
      L3Loop<NX, S, kChunk + kStride, kEnd, kCol0, kAccurate, kRcp, kStride>::run(tD, b3, phase, weight, delta);
    }
  }
};

template <int NX, int S, bool kSplit3, int kRcp, int kTiles>
__global__ void __launch_bounds__(kThreads, 1) rollout_tc2_kernel(Args a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int Lp = NX + 2;
  constexpr int N3t = (2 * NX * S + 15) / 16 * 16;
  constexpr int kChunks = N3t / 16;
  // kTiles == 2: two 128-sample tiles per CTA, 8 warps and 256 TMEM columns each, two column groups per tile, L3 in two
  // column halves.  kTiles == 1 (plans of at most one wave of tiles, where the step latency is all that matters): one
  // tile, all 16 warps on it in four column groups, L3 in one piece (A 128 + D up to 256 columns).
  constexpr int kCG = 4 / kTiles;                                   // column groups per tile
  constexpr int kGroupWarps = 4 * kCG, kGroupT = 32 * kGroupWarps;  // warps / threads per tile
  constexpr int kColsPerThread = kH / kCG, kC16 = kColsPerThread / 16;
  constexpr int kChunksA = (kTiles == 1 || N3t <= 128) ? kChunks : (kChunks + 1) / 2;  // chunks in the first column half
  constexpr int N3a = 16 * kChunksA, N3b = N3t - N3a;
  static_assert((kTiles == 1 || (N3a <= 128 && N3b <= 128)) && N3t <= 256 && Lp + 1 <= 16, "tile shape");
  unsigned char* w1_img = smem_raw;                                   // [hi | lo] 128 x 16 halves = 4 KB each
  unsigned char* w2_img = w1_img + 2 * kH * 16 * 2;                   // [hi | lo] 32 KB each
  unsigned char* w3_img = w2_img + 2 * kH * kH * 2;                   // [hi | lo] N3t * 256 B each
  SmemTail& s = *reinterpret_cast<SmemTail*>(w3_img + 2 * (size_t)N3t * kH * 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  {
    const uint4* s1 = reinterpret_cast<const uint4*>(a.m.mlp2_w1);
    uint4* d1 = reinterpret_cast<uint4*>(w1_img);
    for (int i = tid; i < 2 * kH * 16 * 2 / 16; i += kThreads) d1[i] = __ldg(s1 + i);
    const uint4* s2 = reinterpret_cast<const uint4*>(a.m.mlp2_w2);
    uint4* d2 = reinterpret_cast<uint4*>(w2_img);
    for (int i = tid; i < 2 * kH * kH * 2 / 16; i += kThreads) d2[i] = __ldg(s2 + i);
    const uint4* s3 = reinterpret_cast<const uint4*>(a.m.mlp2_w3);
    uint4* d3 = reinterpret_cast<uint4*>(w3_img);
    for (int i = tid; i < 2 * N3t * kH * 2 / 16; i += kThreads) d3[i] = __ldg(s3 + i);
    for (int i = tid; i < kH; i += kThreads) s.b2[i] = a.m.mlp2_c[i];
    for (int i = tid; i < 256; i += kThreads) s.b3[i] = i < N3t ? a.m.mlp2_c[128 + i] : 0.0f;
    for (int i = tid; i < S; i += kThreads) { s.phase[i] = a.m.ilt_phase[i]; s.weight[i] = a.m.ilt_weight[i]; }
    if (tid < NX) { s.smean[tid] = a.m.state_mean[tid]; s.sinv[tid] = a.m.state_inv_std[tid]; }
    if (tid == 0) {
      for (int g = 0; g < 2; ++g) mbar_init(&s.done[g], 1);
      mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&s.tmem_base, kTmemCols);
  }
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = s.tmem_base;

  // samples of this group: a contiguous range, a multiple of 32 rows long (except at the very end of the plan)
  const int g = warp / kGroupWarps;
  const int n_slots = kTiles * gridDim.x, slot = kTiles * blockIdx.x + g;
  const int per_slot = ((a.K + n_slots - 1) / n_slots + 31) / 32 * 32;
  const long long r_begin_ll = (long long)slot * per_slot;
  const int r_begin = (int)(r_begin_ll < a.K ? r_begin_ll : a.K);
  const int r_end = (r_begin + per_slot < a.K) ? r_begin + per_slot : a.K;
  const uint32_t tg = tmem + kGroupCols * g;
  uint64_t* done = &s.done[g];

  {
    // =====================================  epilogue warps of group g  =====================================
    const int wl = warp % kGroupWarps, q = wl & 3, cg = wl >> 2;
    const int row = 32 * q + lane;
    const uint32_t tlane = tg + ((uint32_t)(32 * q) << 16);
    const uint32_t tA = tlane + kColA, tD = tlane + kColD;
    const uint32_t w1_hi = smem_u32(w1_img), w1_lo = w1_hi + kH * 16 * 2;
    const uint32_t w2_hi = smem_u32(w2_img), w2_lo = w2_hi + kH * kH * 2;
    const uint32_t w3_hi = smem_u32(w3_img), w3_lo = w3_hi + (uint32_t)N3t * kH * 2;
    const uint32_t w3b_off = (uint32_t)(N3a / 8) * kSbo;
    uint32_t n = 0;
    int tstep = 0;
    auto mark = [&](int ev) {
      if (a.trace && blockIdx.x == 0 && lane == 0 && tstep < 104) a.trace[(tstep * 16 + warp) * 8 + ev] = clock64();
    };
    // hand-off to the tensor pipe: every warp of the group has finished its TMEM stores / loads (group barrier), then
    // one thread issues product `ev` of the step.  The group waits for that product anyway (strict chain per sample),
    // so the issuing warp being held by the MMA queue costs nothing; the other group keeps the CUDA cores busy.
    auto issue = [&](int ev) {
      tmem_st_wait();
      fence_before_sync();
      asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"(kGroupT) : "memory");
      if (wl == 0 && lane == 0) {
        fence_after_sync();
        if (ev == 0) issue_gemm_ts<kSplit3>(tg + kColD, tg + kColA, w1_hi, w1_lo, kH, 1, kSbo16);
        else if (ev == 1) issue_gemm_ts<kSplit3>(tg + kColD, tg + kColA, w2_hi, w2_lo, kH, kH / 16, kSbo);
        else if (ev == 2) issue_gemm_ts<kSplit3>(tg + kColD, tg + kColA, w3_hi, w3_lo, N3a, kH / 16, kSbo);
        else issue_gemm_ts<kSplit3>(tg + kColD, tg + kColA, w3_hi + w3b_off, w3_lo + w3b_off, N3b, kH / 16, kSbo);
        mma_commit(done);
      }
      __syncwarp();
    };
    auto wait_mma = [&]() {
      mbar_wait_sleep(done, n & 1); ++n;
      fence_after_sync();
    };
    for (int tile0 = r_begin; tile0 < r_end; tile0 += kRows) {
      const int nrows = (r_end - tile0 < kRows) ? r_end - tile0 : kRows;
      const bool active = 32 * q < nrows;           // warp-uniform: this warp has at least one live sample
      const bool live = row < nrows;
      const int kk = live ? tile0 + row : r_end - 1;
      float st[NX], in[Lp];
#pragma unroll
      for (int c = 0; c < NX; ++c) {
        st[c] = a.state0[(size_t)(a.state_per_sample ? kk / a.state_per_sample : 0) * NX + c];
        in[c] = (st[c] - s.smean[c]) * s.sinv[c];
      }
      {
        const float2 pv = *reinterpret_cast<const float2*>(a.p + ((size_t)kk * a.T) * 2);
        in[NX] = pv.x; in[NX + 1] = pv.y;
      }
      float cost_acc = 0.0f;

      for (int t = 0; t < a.T; ++t) {
        float2 pnext = make_float2(0.f, 0.f);
        if (t + 1 < a.T) pnext = *reinterpret_cast<const float2*>(a.p + ((size_t)kk * a.T + t + 1) * 2);

        // ---------------- A1 = [in | 1 | 0..] as the K = 16 operand (one thread per sample) ----------------
        if (cg == 0 && active) {
          f2_t v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float x0 = (2 * i < Lp) ? in[2 * i < Lp ? 2 * i : 0] : (2 * i == Lp ? 1.0f : 0.0f);
            const float x1 = (2 * i + 1 < Lp) ? in[2 * i + 1 < Lp ? 2 * i + 1 : 0] : (2 * i + 1 == Lp ? 1.0f : 0.0f);
            v[i] = pk2(x0, x1);
          }
          uint32_t ph[8], pl[8];
          pack16<kSplit3>(v, ph, pl);
          tmem_st8(tA, ph);
          if (kSplit3) tmem_st8(tA + 64, pl);
        }
        mark(0);
        issue(0);
        // ---------------- E1: tanh -> A ----------------
        wait_mma();
        mark(1);
        if (active) {
#pragma unroll
          for (int c16 = 0; c16 < kC16; ++c16) {
            const int n0 = kColsPerThread * cg + 16 * c16;
            f2_t v[8];
            ldtm16p(tD + n0, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = tanh2_scaled<kSplit3, kRcp / 10>(v[i]);
            uint32_t ph[8], pl[8];
            pack16<kSplit3>(v, ph, pl);
            tmem_st8(tA + n0 / 2, ph);
            if (kSplit3) tmem_st8(tA + 64 + n0 / 2, pl);
          }
        }
        mark(2);
        issue(1);
        // ---------------- E2: + b2, tanh -> A ----------------
        wait_mma();
        mark(3);
        if (active) {
#pragma unroll
          for (int c16 = 0; c16 < kC16; ++c16) {
            const int n0 = kColsPerThread * cg + 16 * c16;
            f2_t v[8];
            ldtm16p(tD + n0, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 b = *reinterpret_cast<const float4*>(s.b2 + n0 + 4 * i);
              v[2 * i] = tanh2_scaled<kSplit3, kRcp / 10>(add2(v[2 * i], pk2(b.x, b.y)));
              v[2 * i + 1] = tanh2_scaled<kSplit3, kRcp / 10>(add2(v[2 * i + 1], pk2(b.z, b.w)));
            }
            uint32_t ph[8], pl[8];
            pack16<kSplit3>(v, ph, pl);
            tmem_st8(tA + n0 / 2, ph);
            if (kSplit3) tmem_st8(tA + 64 + n0 / 2, pl);
          }
        }
        mark(4);
        issue(2);
        // cost of the previous step's state while the MMA runs (mppi_delay.py:288-290)
        if (cg == 0 && t > 0 && a.cost_total && live)
          cost_acc += env_running_cost_fast(a.o, st, a.hist + ((size_t)kk * a.L + (t - 1) + a.B - 1) * a.nu, a.nu);
        // ---------------- E3: sphere -> complex, Fourier weights, sum over the terms ----------------
        float delta[NX];
#pragma unroll
        for (int c = 0; c < NX; ++c) delta[c] = 0.0f;
        wait_mma();
        mark(5);
        if (active) {
          if (cg == 0) L3Loop<NX, S, 0, kChunksA, 0, kSplit3, kRcp % 10, kCG>::run(tD, s.b3, s.phase, s.weight, delta);
          else if (cg == 1) L3Loop<NX, S, 1, kChunksA, 0, kSplit3, kRcp % 10, kCG>::run(tD, s.b3, s.phase, s.weight, delta);
          else if (kCG == 4 && cg == 2) L3Loop<NX, S, 2, kChunksA, 0, kSplit3, kRcp % 10, kCG>::run(tD, s.b3, s.phase, s.weight, delta);
          else if (kCG == 4) L3Loop<NX, S, 3, kChunksA, 0, kSplit3, kRcp % 10, kCG>::run(tD, s.b3, s.phase, s.weight, delta);
        }
        if (N3b > 0) {
          mark(6);
          issue(3);
          wait_mma();
          mark(7);
          if (active) {
            constexpr int kFirstB0 = kChunksA + (kChunksA & 1);        // first chunk >= kChunksA with even index
            constexpr int kFirstB1 = kChunksA + 1 - (kChunksA & 1);    // ... with odd index
            if (cg == 0) L3Loop<NX, S, kFirstB0, kChunks, kChunksA, kSplit3, kRcp % 10, 2>::run(tD, s.b3, s.phase, s.weight, delta);
            else L3Loop<NX, S, kFirstB1, kChunks, kChunksA, kSplit3, kRcp % 10, 2>::run(tD, s.b3, s.phase, s.weight, delta);
          }
        }
#pragma unroll
        for (int c = 0; c < NX; ++c) s.exch[g * kCG + cg][row * NX + c] = delta[c];
        asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"(kGroupT) : "memory");
        // every thread of a sample adds the partials in the same order: the replicated state stays bit-identical
#pragma unroll
        for (int c = 0; c < NX; ++c) {
          float d = s.exch[g * kCG][row * NX + c];
#pragma unroll
          for (int j = 1; j < kCG; ++j) d += s.exch[g * kCG + j][row * NX + c];
          st[c] += d;                                             // mppi_with_model.py:121
          in[c] = (st[c] - s.smean[c]) * s.sinv[c];
          if (cg == kCG - 1 && live) {
            if (a.states) a.states[((size_t)kk * a.T + t) * NX + c] = st[c];
            if (a.delta_out) a.delta_out[(size_t)kk * NX + c] = d;
          }
        }
        in[NX] = pnext.x; in[NX + 1] = pnext.y;
        ++tstep;
        // exch is rewritten only after the next step's MMA round trips, which need every warp of the group
      }
      if (cg == 0 && live && a.cost_total) {
        cost_acc += env_running_cost_fast(a.o, st, a.hist + ((size_t)kk * a.L + (a.T - 1) + a.B - 1) * a.nu, a.nu);
        a.cost_total[kk] = cost_acc + (a.pert_cost ? a.pert_cost[kk] : 0.0f);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

template <int NX, int S, bool kSplit3, int kRcp, int kTiles>
static int launch_one_t(const Args& a, cudaStream_t stream) {
  constexpr int N3t = (2 * NX * S + 15) / 16 * 16;
  const size_t smem = 2 * (size_t)kH * 16 * 2 + 2 * (size_t)kH * kH * 2 + 2 * (size_t)N3t * kH * 2 + sizeof(SmemTail) + 128;
  NLC_REQUIRE(smem <= 227 * 1024, NLC_ERR_SHAPE, "tcgen05 rollout: %zu bytes of shared memory needed", smem);
  auto kern = rollout_tc2_kernel<NX, S, kSplit3, kRcp, kTiles>;
  NLC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // one tile slot per 32..128 samples: enough CTAs to give every slot at least one warp of samples, at most one per SM
  int grid = (a.K + 32 * kTiles - 1) / (32 * kTiles);
  if (grid > 148) grid = 148;
  kern<<<grid, kThreads, smem, stream>>>(a);
  NLC_LAUNCH_OK("rollout_tc2_kernel");
  return NLC_OK;
}

// two tiles per CTA once the plan is more than one wave of 128-sample tiles, else one tile on all 16 warps
template <int NX, int S, bool kSplit3, int kRcp>
static int launch_one(const Args& a, int tiles, cudaStream_t stream) {
  return tiles == 1 ? launch_one_t<NX, S, kSplit3, kRcp, 1>(a, stream) : launch_one_t<NX, S, kSplit3, kRcp, 2>(a, stream);
}

}  // namespace rt2

// returns NLC_ERR_UNSUPPORTED when the (nx, S) pair has no tensor-core instantiation (caller falls back)
static long long* g_roll_trace = nullptr;
void set_rollout_trace(long long* p) { g_roll_trace = p; }

int launch_rollout_tc2(nlc_model_s* m, const nlc_rollout_opts* o, const float* state, int sps, const float* p,
                       const float* hist, const float* pert_cost, int K, int T, int B, int nu, float* cost, float* states,
                       float* delta_out, int split3, int tiles_per_cta, cudaStream_t stream) {
  rt2::Args a;
  a.m = m->d; a.o = *o; a.state0 = state; a.state_per_sample = sps; a.p = p; a.hist = hist; a.pert_cost = pert_cost;
  a.K = K; a.T = T; a.B = B; a.L = B - 1 + T; a.nu = nu; a.cost_total = cost; a.states = states; a.delta_out = delta_out;
  a.trace = g_roll_trace;
  // reciprocal flavour of the fp32-class epilogues: 3 = Newton on the FMA pipe (default: 1.71 ms at config 4),
  // 0 = MUFU.RCP (1.85 ms: shorter chains, but the MUFU pipe is the scarcer one).  NLC_ROLLOUT_RCP is a measurement knob.
  static const int rcp = [] { const char* e = getenv("NLC_ROLLOUT_RCP"); return (e && e[0] && e[1]) ? (e[0] - '0') * 10 + (e[1] - '0') : 33; }();
#define NLC_RT2_CASE(NX_, S_)                                                                                   \
  if (m->nx == NX_ && m->S == S_) {                                                                             \
    if (!split3) return rt2::launch_one<NX_, S_, false, 0>(a, tiles_per_cta, stream);  /* digits: (MLP tanh, L3 pair) reciprocal flavour */                                          \
    if (rcp == 0) return rt2::launch_one<NX_, S_, true, 0>(a, tiles_per_cta, stream);                                          \
    if (rcp == 3) return rt2::launch_one<NX_, S_, true, 3>(a, tiles_per_cta, stream);                                          \
    if (rcp == 30) return rt2::launch_one<NX_, S_, true, 30>(a, tiles_per_cta, stream);                                        \
    return rt2::launch_one<NX_, S_, true, 33>(a, tiles_per_cta, stream);                                                       \
  }
  NLC_RT2_CASE(3, 17)
  NLC_RT2_CASE(5, 17)
  NLC_RT2_CASE(6, 17)
  NLC_RT2_CASE(3, 33)
#undef NLC_RT2_CASE
  set_error("tcgen05 rollout has no instantiation for nx=%d S=%d", m->nx, m->S);
  return NLC_ERR_UNSUPPORTED;
}

}  // namespace nlc

// measurement hook (tools/trace_rollout.py): device buffer of 32*16*8 int64 receiving CTA 0's clock64 timeline
extern "C" void nlc_debug_set_rollout_trace(void* dev_ptr) { nlc::set_rollout_trace(static_cast<long long*>(dev_ptr)); }
