// Stages 2+3 for the Neural Laplace dynamics with the representation MLP on the 5th-generation tensor cores.
//
// Same recurrence as rollout.cu (state <- state + ILT(MLP([s | obs_n | p_action])), cost += running_cost), one
// CTA = 128 samples for the whole horizon, one launch per plan.  Per step:
//
//   L1  (nx+2 -> 128, the 2S constant s-columns folded into the bias) on the CUDA cores, written straight into
//       TENSOR MEMORY as the fp16 hi/lo A operand of the next product (tcgen05.st; no shared-memory round trip);
//   L2  128 -> 128      tcgen05.mma kind::f16, A from TMEM, W2 image resident in shared memory, fp32 accumulate;
//       epilogue (bias + tanh) re-writes the A region with the hidden activations;
//   L3  128 -> 2*nx*S   tcgen05.mma, W3 image (rows pair-permuted: theta_k, phi_k adjacent) resident in shared memory;
//       epilogue applies the sphere->complex map and the Fourier weights per (channel, term) pair and sums over the
//       terms in a fixed order; the two column groups of a sample exchange their partial sums through shared memory.
//
//   NLC_MATH_TC_SPLIT3: operands split hi+lo in fp16, D += A_hi B_hi + A_lo B_hi + A_hi B_lo (fp32-class);
//   NLC_MATH_TC_FP16  : single pass (11-bit operands; the stated looser bound).
//
// TMEM: A region 128 columns (hi 64 | lo 64, two fp16 per column), accumulator region N3t <= 208 columns shared by
// L2 and L3 (never live together).  Shared memory: only the weight images (<= 170 KB) and small constants.
// Threads: kGroups*128; warp w owns TMEM lanes 32(w&3).. (its 32 samples) and column group w>>2.  The recurrence is a
// strict chain per sample, so MMA and epilogue alternate; each thread keeps its sample's state in registers.
#include <cuda_fp16.h>

#include "common.cuh"
#include "env_cost.cuh"
#include "tc_umma.cuh"

namespace nlc {

using namespace umma;

constexpr int kRtRows = 128;
constexpr int kRtH = 128;
constexpr uint32_t kRtLbo = 128, kRtSbo = (kRtH / 8) * 128;  // K = 128
constexpr uint32_t kRtColA = 0, kRtColD = 128, kRtTmemCols = 512;

struct RollTcArgs {
  ModelDev m;
  nlc_rollout_opts o;
  const float* state0; int state_per_sample;
  const float* p;
  const float* hist;
  const float* pert_cost;
  int K, T, B, L, nu, N3t;
  int rows_per_cta;  // 128, or 64 when K is too small to fill the SMs with full tiles (upper TMEM lanes idle)
  float* cost_total;
  float* states;
  float* delta_out;
};

__device__ __forceinline__ float rt_sigmoid_fast(float x) {
  float e = ex2_approx(-1.44269504088896f * x);
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}
__device__ __forceinline__ float rt_tanh_fast(float x) { return fmaf(2.0f, rt_sigmoid_fast(2.0f * x), -1.0f); }

// tan(phi/2 + pi/4) with phi = (pi/2) tanh(u) = tan((pi/2) sigmoid(2u)), through s = sigmoid(-2|u|) in (0, 1/2] (see
// common.cuh:sphere_radius for the derivation); MUFU ex2/rcp directly: relative error <= ~1.5e-6 for |u| <= 8.
__device__ __forceinline__ float rt_sphere_radius(float u) {
  const float e = ex2_approx(-2.88539008177793f * fabsf(u));
  float inv;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(1.0f + e));
  const float x = 1.57079632679489662f * (e * inv);  // (0, pi/4]
  const float x2 = x * x;
  const float sn = x * fmaf(x2, fmaf(x2, fmaf(x2, fmaf(x2, 2.7557319224e-6f, -1.9841269841e-4f), 8.3333333333e-3f), -1.6666666667e-1f), 1.0f);
  const float cs = fmaf(x2, fmaf(x2, fmaf(x2, fmaf(x2, 2.4801587302e-5f, -1.3888888889e-3f), 4.1666666667e-2f), -0.5f), 1.0f);
  const float num = u <= 0.0f ? sn : cs, den = u <= 0.0f ? cs : sn;
  float dinv;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(dinv) : "f"(den));
  return num * dinv;
}
// cos(pi * (y + c)) for y in (-1, 1), c in (-1, 1]: exact reduction of the argument to [-1, 1] half-turns, then
// cos.approx (abs error 2^-21.4 on [-pi, pi]).
__device__ __forceinline__ float rt_cospi_sum(float y, float c) {
  float z = y + c;
  z = fmaf(-2.0f, rintf(0.5f * z), z);
  return __cosf(3.14159265358979f * z);
}

// 16 fp32 values -> 8 packed fp16 pairs (hi) and the fp16 residuals (lo)
template <bool kSplit3>
__device__ __forceinline__ void pack16(const float (&v)[16], uint32_t (&ph)[8], uint32_t (&pl)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const __half h0 = __float2half_rn(v[2 * i]), h1 = __float2half_rn(v[2 * i + 1]);
    const __half2 hh = __halves2half2(h0, h1);
    ph[i] = *reinterpret_cast<const uint32_t*>(&hh);
    if (kSplit3) {
      const __half2 ll = __halves2half2(__float2half_rn(v[2 * i] - __half2float(h0)), __float2half_rn(v[2 * i + 1] - __half2float(h1)));
      pl[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
  }
}

// D[128 x N] = A[128 x 128] (TMEM) * B[N x 128]^T (smem image)
template <bool kSplit3>
__device__ __forceinline__ void issue_gemm_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_hi, uint32_t b_lo, int N) {
  const uint32_t idesc = idesc_f16_f32(kRtRows, N);
#pragma unroll
  for (int ks = 0; ks < kRtH / 16; ++ks) {
    const uint32_t boff = ks * 2 * kRtLbo;
    mma_f16_ts(d_tmem, a_tmem + 8 * ks, smem_desc(b_hi + boff, kRtLbo, kRtSbo), idesc, ks > 0 ? 1u : 0u);
    if (kSplit3) {
      mma_f16_ts(d_tmem, a_tmem + 64 + 8 * ks, smem_desc(b_hi + boff, kRtLbo, kRtSbo), idesc, 1u);
      mma_f16_ts(d_tmem, a_tmem + 8 * ks, smem_desc(b_lo + boff, kRtLbo, kRtSbo), idesc, 1u);
    }
  }
}

// L3 epilogue of one 16-column chunk (8 (channel, term) pairs) for compile-time chunk index kChunk
template <int NX, int S, int kChunk>
__device__ __forceinline__ void l3_chunk(uint32_t tlane, const float* __restrict__ b3, const float* __restrict__ phase,
                                         const float* __restrict__ weight, float (&delta)[NX]) {
  float v[16];
  tmem_ld16(tlane + kRtColD + 16 * kChunk, v);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    constexpr int dummy = 0; (void)dummy;
    const int pair = 8 * kChunk + i;
    if (pair < NX * S) {
      const int ch = pair / S, k = pair - ch * S;
      const float y = rt_tanh_fast(v[2 * i] + b3[2 * pair]);                       // theta = pi y, w_nl.py:59
      const float rad = rt_sphere_radius(v[2 * i + 1] + b3[2 * pair + 1]);         // w_nl.py:60-62 + sphere_to_complex
      delta[ch] = fmaf(weight[k] * rad, rt_cospi_sum(y, phase[k]), delta[ch]);     // phase[] holds k t/T in half-turns
    }
  }
}

template <int NX, int S, int kGroups, int kGroup, int kChunk>
struct L3Loop {
  static __device__ __forceinline__ void run(uint32_t tlane, const float* b3, const float* phase, const float* weight, float (&delta)[NX]) {
    constexpr int kNChunk = (2 * NX * S + 15) / 16;
    if constexpr (kChunk < kNChunk) {
      l3_chunk<NX, S, kChunk>(tlane, b3, phase, weight, delta);
      L3Loop<NX, S, kGroups, kGroup, kChunk + kGroups>::run(tlane, b3, phase, weight, delta);
    }
  }
};

struct RollTcSmemTail {  // after the weight images
  float w1x[(kMaxNx + 2) * kRtH];
  float b1[kRtH], b2[kRtH];
  float b3[256];
  float phase[kMaxS], weight[kMaxS];
  float smean[kMaxNx], sinv[kMaxNx];
  float exch[4][kRtRows * kMaxNx];  // per column group partial ILT sums
  alignas(8) uint64_t bar;
  uint32_t tmem_base;
};

template <int NX, int S, int kGroups, bool kSplit3>
__global__ void __launch_bounds__(kGroups * 128, 1) rollout_tc_kernel(RollTcArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int kThreads = kGroups * 128;
  constexpr int Lp = NX + 2;
  constexpr int kColsPerGroup = kRtH / kGroups;  // L1 / L2 columns per thread
  const int N3t = a.N3t;
  unsigned char* w2_img = smem_raw;                                   // [hi | lo] 32 KB each
  unsigned char* w3_img = w2_img + 2 * kRtH * kRtH * 2;               // [hi | lo] N3t*256 B each
  RollTcSmemTail& s = *reinterpret_cast<RollTcSmemTail*>(w3_img + 2 * (size_t)N3t * kRtH * 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, grp = warp >> 2;
  const int row = 32 * q + lane;
  const int k0 = blockIdx.x * a.rows_per_cta;
  const int kk = min(k0 + row, a.K - 1);
  const bool live = row < a.rows_per_cta && k0 + row < a.K;

  {
    const uint4* s2 = reinterpret_cast<const uint4*>(a.m.mlp_tc_w2);
    uint4* d2 = reinterpret_cast<uint4*>(w2_img);
    for (int i = tid; i < 2 * kRtH * kRtH * 2 / 16; i += kThreads) d2[i] = __ldg(s2 + i);
    const uint4* s3 = reinterpret_cast<const uint4*>(a.m.mlp_tc_w3);
    uint4* d3 = reinterpret_cast<uint4*>(w3_img);
    for (int i = tid; i < 2 * N3t * kRtH * 2 / 16; i += kThreads) d3[i] = __ldg(s3 + i);
    for (int i = tid; i < Lp * kRtH; i += kThreads) s.w1x[i] = a.m.w1x_t[i];
    for (int i = tid; i < kRtH; i += kThreads) { s.b1[i] = a.m.b1_fold[i]; s.b2[i] = a.m.b2[i]; }
    for (int i = tid; i < N3t; i += kThreads) s.b3[i] = a.m.b3_tc[i];
    for (int i = tid; i < S; i += kThreads) { s.phase[i] = a.m.ilt_phase[i] * 0.318309886183791f; s.weight[i] = a.m.ilt_weight[i]; }
    if (tid < NX) { s.smean[tid] = a.m.state_mean[tid]; s.sinv[tid] = a.m.state_inv_std[tid]; }
    if (tid == 0) { mbar_init(&s.bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&s.tmem_base, kRtTmemCols);
  }
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = s.tmem_base;
  const uint32_t tlane = tmem + ((uint32_t)(32 * q) << 16);
  const uint32_t w2_hi = smem_u32(w2_img), w2_lo = w2_hi + kRtH * kRtH * 2;
  const uint32_t w3_hi = smem_u32(w3_img), w3_lo = w3_hi + (uint32_t)N3t * kRtH * 2;
  uint32_t par = 0;

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
    // prefetch the next step's encoder output (hidden behind the three phases below)
    float2 pnext = make_float2(0.f, 0.f);
    if (t + 1 < a.T) pnext = *reinterpret_cast<const float2*>(a.p + ((size_t)kk * a.T + t + 1) * 2);

    // ---------------- L1 on the CUDA cores -> A region (TMEM) ----------------
#pragma unroll
    for (int c16 = 0; c16 < kColsPerGroup / 16; ++c16) {
      const int n0 = grp * kColsPerGroup + 16 * c16;
      float v[16];
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {
        float4 acc = *reinterpret_cast<const float4*>(s.b1 + n0 + 4 * i4);
#pragma unroll
        for (int j = 0; j < Lp; ++j) {
          const float4 w = *reinterpret_cast<const float4*>(s.w1x + j * kRtH + n0 + 4 * i4);
          acc.x = fmaf(w.x, in[j], acc.x); acc.y = fmaf(w.y, in[j], acc.y);
          acc.z = fmaf(w.z, in[j], acc.z); acc.w = fmaf(w.w, in[j], acc.w);
        }
        v[4 * i4] = rt_tanh_fast(acc.x); v[4 * i4 + 1] = rt_tanh_fast(acc.y);
        v[4 * i4 + 2] = rt_tanh_fast(acc.z); v[4 * i4 + 3] = rt_tanh_fast(acc.w);
      }
      uint32_t ph[8], pl[8];
      pack16<kSplit3>(v, ph, pl);
      tmem_st8(tlane + kRtColA + n0 / 2, ph);
      if (kSplit3) tmem_st8(tlane + kRtColA + 64 + n0 / 2, pl);
    }
    tmem_st_wait();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      issue_gemm_ts<kSplit3>(tmem + kRtColD, tmem + kRtColA, w2_hi, w2_lo, kRtH);
      mma_commit(&s.bar);
    }
    // ---------------- L2 epilogue: bias + tanh -> A region ----------------
    mbar_wait(&s.bar, par); par ^= 1;
    fence_after_sync();
#pragma unroll
    for (int c16 = 0; c16 < kColsPerGroup / 16; ++c16) {
      const int n0 = grp * kColsPerGroup + 16 * c16;
      float v[16];
      tmem_ld16(tlane + kRtColD + n0, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = rt_tanh_fast(v[i] + s.b2[n0 + i]);
      uint32_t ph[8], pl[8];
      pack16<kSplit3>(v, ph, pl);
      tmem_st8(tlane + kRtColA + n0 / 2, ph);
      if (kSplit3) tmem_st8(tlane + kRtColA + 64 + n0 / 2, pl);
    }
    tmem_st_wait();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      issue_gemm_ts<kSplit3>(tmem + kRtColD, tmem + kRtColA, w3_hi, w3_lo, N3t);
      mma_commit(&s.bar);
    }
    // cost of the previous step's state while the MMA runs (mppi_delay.py:288-290)
    if (grp == 0 && t > 0 && a.cost_total)
      cost_acc += env_running_cost(a.o, st, a.hist + ((size_t)kk * a.L + (t - 1) + a.B - 1) * a.nu, a.nu);
    // ---------------- L3 epilogue: sphere -> complex, Fourier weights, sum over the terms ----------------
    mbar_wait(&s.bar, par); par ^= 1;
    fence_after_sync();
    float delta[NX];
#pragma unroll
    for (int c = 0; c < NX; ++c) delta[c] = 0.0f;
    if (kGroups == 2) {
      if (grp == 0) L3Loop<NX, S, 2, 0, 0>::run(tlane, s.b3, s.phase, s.weight, delta);
      else L3Loop<NX, S, 2, 1, 1>::run(tlane, s.b3, s.phase, s.weight, delta);
    } else {
      if (grp == 0) L3Loop<NX, S, 4, 0, 0>::run(tlane, s.b3, s.phase, s.weight, delta);
      else if (grp == 1) L3Loop<NX, S, 4, 1, 1>::run(tlane, s.b3, s.phase, s.weight, delta);
      else if (grp == 2) L3Loop<NX, S, 4, 2, 2>::run(tlane, s.b3, s.phase, s.weight, delta);
      else L3Loop<NX, S, 4, 3, 3>::run(tlane, s.b3, s.phase, s.weight, delta);
    }
#pragma unroll
    for (int c = 0; c < NX; ++c) s.exch[grp][row * NX + c] = delta[c];
    fence_before_sync();
    __syncthreads();
    // every thread of a sample adds the partials in the same order: the replicated state stays bit-identical
#pragma unroll
    for (int c = 0; c < NX; ++c) {
      float d = s.exch[0][row * NX + c];
#pragma unroll
      for (int g = 1; g < kGroups; ++g) d += s.exch[g][row * NX + c];
      st[c] += d;                                             // mppi_with_model.py:121
      in[c] = (st[c] - s.smean[c]) * s.sinv[c];
      if (grp == kGroups - 1 && live) {
        if (a.states) a.states[((size_t)(k0 + row) * a.T + t) * NX + c] = st[c];
        if (a.delta_out) a.delta_out[(size_t)(k0 + row) * NX + c] = d;
      }
    }
    in[NX] = pnext.x; in[NX + 1] = pnext.y;
    // exch is rewritten only after the next step's two barriers
  }
  if (grp == 0 && live && a.cost_total) {
    cost_acc += env_running_cost(a.o, st, a.hist + ((size_t)kk * a.L + (a.T - 1) + a.B - 1) * a.nu, a.nu);
    a.cost_total[k0 + row] = cost_acc + (a.pert_cost ? a.pert_cost[k0 + row] : 0.0f);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kRtTmemCols);
}

template <int NX, int S, int kGroups, bool kSplit3>
static int launch_one(const RollTcArgs& a, cudaStream_t stream) {
  const size_t smem = 2 * (size_t)kRtH * kRtH * 2 + 2 * (size_t)a.N3t * kRtH * 2 + sizeof(RollTcSmemTail) + 128;
  NLC_REQUIRE(smem <= 227 * 1024, NLC_ERR_SHAPE, "tcgen05 rollout: %zu bytes of shared memory needed", smem);
  auto kern = rollout_tc_kernel<NX, S, kGroups, kSplit3>;
  NLC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (a.K + a.rows_per_cta - 1) / a.rows_per_cta;
  kern<<<grid, kGroups * 128, smem, stream>>>(a);
  NLC_LAUNCH_OK("rollout_tc_kernel");
  return NLC_OK;
}

// returns NLC_ERR_UNSUPPORTED when the (nx, S) pair has no tensor-core instantiation (caller falls back to FFMA)
int launch_rollout_tc(nlc_model_s* m, const nlc_rollout_opts* o, const float* state, int sps, const float* p,
                      const float* hist, const float* pert_cost, int K, int T, int B, int nu, float* cost, float* states,
                      float* delta_out, int split3, int groups, cudaStream_t stream) {
  RollTcArgs a;
  a.m = m->d; a.o = *o; a.state0 = state; a.state_per_sample = sps; a.p = p; a.hist = hist; a.pert_cost = pert_cost;
  // the horizon is sequential per tile, so a plan with fewer than ~one wave of 128-sample tiles runs faster on twice as
  // many half-filled tiles (the MMA costs the same, the epilogue halves)
  a.rows_per_cta = ((K + kRtRows - 1) / kRtRows <= 74) ? 64 : kRtRows;
  a.K = K; a.T = T; a.B = B; a.L = B - 1 + T; a.nu = nu; a.N3t = m->N3t; a.cost_total = cost; a.states = states; a.delta_out = delta_out;
#define NLC_RT_CASE(NX_, S_)                                                                         \
  if (m->nx == NX_ && m->S == S_) {                                                                  \
    if (groups == 4) return split3 ? launch_one<NX_, S_, 4, true>(a, stream) : launch_one<NX_, S_, 4, false>(a, stream); \
    return split3 ? launch_one<NX_, S_, 2, true>(a, stream) : launch_one<NX_, S_, 2, false>(a, stream); \
  }
  NLC_RT_CASE(3, 17)
  NLC_RT_CASE(5, 17)
  NLC_RT_CASE(6, 17)
  NLC_RT_CASE(3, 33)
#undef NLC_RT_CASE
  set_error("tcgen05 rollout has no instantiation for nx=%d S=%d", m->nx, m->S);
  return NLC_ERR_UNSUPPORTED;
}

}  // namespace nlc
