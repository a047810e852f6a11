// Delayed-history encoder (ReverseGRUEncoder.forward, w_nl.py:25-29) on the 5th-generation tensor cores.
//
// Same dataflow as encode_gru.cu (all K*T windows of a plan in one wide pass, zero-state products skipped), with
// the ten 64x192 recurrent products of every window issued as tcgen05.mma (kind::f16, M=128 windows per CTA,
// fp32 accumulators in TMEM) and only the gate nonlinearities on the CUDA cores:
//
//   operands   fp16 hi (+ fp16 lo = fp16(x - hi)) images in shared memory, K-major no-swizzle canonical layout.
//              NLC_MATH_TC_SPLIT3: D += A_hi B_hi + A_lo B_hi + A_hi B_lo  (22 significand bits per operand, fp32
//              accumulate: fp32-class results, 3 MMAs);  NLC_MATH_TC_FP16: D += A_hi B_hi (11 bits, looser bound).
//   weights    the three recurrent matrices (hi and lo, 144 KB) stay resident in shared memory for the CTA's
//              persistent loop; hidden states h0/h1 are re-written as A operands by the epilogue (64 KB).
//   TMEM       D0[192] layer-0 hidden pre-activations (r,z,n) | D1_rz[128] layer-1 r,z (input + hidden products
//              accumulate into the same columns) | D1_in[64] | D1_hn[64]  = 448 of 512 columns.
//   schedule   one thread issues; the layer-0 chain (A cells) runs one cell ahead of the layer-1 chain (B cells) so
//              that every MMA burst overlaps the other chain's gate epilogue:
//                 epi A(s+1) || MMA B(s)   ->   epi B(s) || MMA A(s+2)   ->   ...
//              Completion is tracked with tcgen05.commit on two mbarriers; shared-memory operands are single
//              buffered (a state is overwritten only after the commit of every MMA that reads it).
//   threads    16 warps; warp w owns TMEM lanes 32(w&3).. (its 32 windows) and hidden units 16(w>>2)..; each thread
//              keeps its window's 16+16 fp32 hidden values in registers, so the fp16 images are write-only.
//              (The epilogue is MUFU/issue bound, not MMA bound: 16 warps double the latency hiding of 8.)
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_pack.cuh"
#include "tc_umma.cuh"

namespace nlc {

using namespace umma;

constexpr int kTcRows = 128;
constexpr int kTcHg = 64;
constexpr int kTcG3 = 192;
constexpr int kTcThreads = 512;
constexpr int kTcUnits = 16;  // hidden units per thread
constexpr uint32_t kOpBytes = kTcRows * kTcHg * 2;  // one fp16 A-operand image (16 KB)
constexpr uint32_t kWBytes = kTcG3 * kTcHg * 2;     // one fp16 weight image (24 KB)
constexpr uint32_t kLbo = 128, kSbo = (kTcHg / 8) * 128;  // K-major no-swizzle, K = 64
constexpr uint32_t kColD0 = 0, kColRz = 192, kColIn = 320, kColHn = 384, kTmemCols = 512;

struct EncTcArgs {
  const float* hist;
  float* p_out;
  int K, T, B, L, gin, hist_ch;
  long long rows;
  ModelDev m;
};

struct EncTcSmem {
  alignas(128) unsigned char w[3][2][kWBytes];  // [W_hh0, W_ih1, W_hh1][hi, lo]
  alignas(128) unsigned char h0[2][kOpBytes];   // [hi, lo]
  alignas(128) unsigned char h1[2][kOpBytes];
  alignas(16) float brz0[128], bin0[64], bhn0[64], brz1[128], bin1[64], bhn1[64];
  alignas(16) float w_ih0[3 * kMaxNu * kTcHg];  // [gate][input v][unit]: contiguous over units for 16-byte loads
  alignas(16) float w_out[2 * kTcHg];
  alignas(16) float act[2][kTcRows * 8];  // double-buffered [row][B*gin], B*gin <= 8
  alignas(16) float pout[3][kTcRows * 2];
  float b_out[2];
  float act_mean[kMaxNu], act_inv_std[kMaxNu];
  alignas(8) uint64_t bar_a, bar_b;
  uint32_t tmem_base;
};

__device__ __forceinline__ float sigmoid_fast(float x) {  // abs error <= ~3e-7 for |x| <= 16
  float e = ex2_approx(-1.44269504088896f * x);
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}
__device__ __forceinline__ float tanh_fast(float x) { return fmaf(2.0f, sigmoid_fast(2.0f * x), -1.0f); }
__device__ __forceinline__ float tanh_mufu(float x) {  // one MUFU, |error| <= 2^-10.99: single-pass fp16 mode only
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// GRU cell update for one hidden unit from the biased pre-activations.  kAccurate: 2 ex2 + 1 shared rcp for (r, z),
// ex2 + rcp for n (5 MUFU, abs error ~3e-7); otherwise 3 tanh.approx (the fp16 single-pass mode's accuracy class).
template <bool kAccurate>
__device__ __forceinline__ float gru_unit(float pre_r, float pre_z, float gi_n, float gh_n, float h_old) {
  float r, z, n;
  if (kAccurate) {
    const float ea = ex2_approx(-1.44269504088896f * pre_r), eb = ex2_approx(-1.44269504088896f * pre_z);
    const float da = 1.0f + ea, db = 1.0f + eb;
    float inv;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(da * db));
    r = db * inv;
    z = da * inv;
    n = tanh_fast(fmaf(r, gh_n, gi_n));
  } else {
    r = fmaf(0.5f, tanh_mufu(0.5f * pre_r), 0.5f);
    z = fmaf(0.5f, tanh_mufu(0.5f * pre_z), 0.5f);
    n = tanh_mufu(fmaf(r, gh_n, gi_n));
  }
  return fmaf(z, h_old - n, n);  // (1 - z) n + z h
}

__device__ __forceinline__ void lds16(const float* p, float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = *reinterpret_cast<const float4*>(p + 4 * i);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
__device__ __forceinline__ void lds8(const float* p, float (&v)[8]) {
  const float4 t0 = *reinterpret_cast<const float4*>(p), t1 = *reinterpret_cast<const float4*>(p + 4);
  v[0] = t0.x; v[1] = t0.y; v[2] = t0.z; v[3] = t0.w; v[4] = t1.x; v[5] = t1.y; v[6] = t1.z; v[7] = t1.w;
}
__device__ __forceinline__ void bias_to_tmem8(uint32_t taddr, const float* b8) {
  float v[8];
  lds8(b8, v);
  uint32_t r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(v[i]);
  tmem_st8(taddr, r);
}
// (re-)initialise 16 accumulator columns of this thread's TMEM lane with a bias vector: the next cell's MMAs
// accumulate on top of it, so no bias is loaded or added in the gate epilogue
__device__ __forceinline__ void bias_to_tmem(uint32_t taddr, const float* b16) {
  float v[16];
  lds16(b16, v);
  uint32_t r[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(v[i]);
  tmem_st16(taddr, r);
}

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

// this thread's 32 hidden values -> fp16 hi (+lo) A-operand image rows
template <bool kSplit3>
__device__ __forceinline__ void store_operand(unsigned char* img_hi, unsigned char* img_lo, int row, int unit0, const float (&h)[kTcUnits]) {
#pragma unroll
  for (int g8 = 0; g8 < kTcUnits / 8; ++g8) {
    uint32_t ph[4], pl[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float x0 = h[8 * g8 + 2 * i], x1 = h[8 * g8 + 2 * i + 1];
      const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
      const __half2 hh = __halves2half2(h0, h1);
      ph[i] = *reinterpret_cast<const uint32_t*>(&hh);
      if (kSplit3) {
        const __half2 ll = __halves2half2(__float2half_rn(x0 - __half2float(h0)), __float2half_rn(x1 - __half2float(h1)));
        pl[i] = *reinterpret_cast<const uint32_t*>(&ll);
      }
    }
    const uint32_t off = (uint32_t)(row >> 3) * kSbo + (uint32_t)((unit0 >> 3) + g8) * kLbo + (uint32_t)(row & 7) * 16;
    *reinterpret_cast<uint4*>(img_hi + off) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    if (kSplit3) *reinterpret_cast<uint4*>(img_lo + off) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
  }
}

template <bool kSplit3, int GIN>
__global__ void __launch_bounds__(kTcThreads, 1) encode_tc_kernel(EncTcArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  EncTcSmem& s = *reinterpret_cast<EncTcSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, grp = warp >> 2;
  const int row = 32 * q + lane;       // window within the tile == TMEM lane
  const int unit0 = kTcUnits * grp;    // first hidden unit this thread owns
  const int B = a.B, BG = B * GIN;
  constexpr int gin = GIN;

  // ---- one-time setup: weight images, biases, barriers, TMEM ----
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.m.enc_tc_w);
    uint4* dst = reinterpret_cast<uint4*>(&s.w[0][0][0]);
    for (int i = tid; i < (int)(3 * 2 * kWBytes / 16); i += kTcThreads) dst[i] = __ldg(src + i);
    for (int i = tid; i < 128; i += kTcThreads) {
      s.brz0[i] = a.m.b_ih0[i] + a.m.b_hh0[i];
      s.brz1[i] = a.m.b_ih1[i] + a.m.b_hh1[i];
    }
    for (int i = tid; i < 64; i += kTcThreads) {
      s.bin0[i] = a.m.b_ih0[128 + i]; s.bhn0[i] = a.m.b_hh0[128 + i];
      s.bin1[i] = a.m.b_ih1[128 + i]; s.bhn1[i] = a.m.b_hh1[128 + i];
    }
    for (int i = tid; i < kTcG3 * gin; i += kTcThreads) {  // global [3*64][gin] -> [gate][v][unit]
      const int gu = i / gin, v = i - gu * gin, g = gu >> 6, u = gu & 63;
      s.w_ih0[(g * kMaxNu + v) * kTcHg + u] = a.m.w_ih0[i];
    }
    for (int i = tid; i < 2 * kTcHg; i += kTcThreads) s.w_out[i] = a.m.w_out[i];
    if (tid < 2) s.b_out[tid] = a.m.b_out[tid];
    if (tid < gin) { s.act_mean[tid] = a.m.act_mean[tid]; s.act_inv_std[tid] = a.m.act_inv_std[tid]; }
    if (tid == 0) { mbar_init(&s.bar_a, 1); mbar_init(&s.bar_b, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&s.tmem_base, kTmemCols);
  }
  fence_proxy_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = s.tmem_base;
  const uint32_t tlane = tmem + ((uint32_t)(32 * q) << 16);
  const uint32_t a_h0_hi = smem_u32(s.h0[0]), a_h0_lo = smem_u32(s.h0[1]);
  const uint32_t a_h1_hi = smem_u32(s.h1[0]), a_h1_lo = smem_u32(s.h1[1]);
  const uint32_t w_hh0_hi = smem_u32(s.w[0][0]), w_hh0_lo = smem_u32(s.w[0][1]);
  const uint32_t w_ih1_hi = smem_u32(s.w[1][0]), w_ih1_lo = smem_u32(s.w[1][1]);
  const uint32_t w_hh1_hi = smem_u32(s.w[2][0]), w_hh1_lo = smem_u32(s.w[2][1]);
  const uint32_t n_rows_off = (128 / 8) * kSbo;  // weight rows 128..191 (the n gate)

  // accumulator columns start out holding the biases (and are reset to them after every read, see bias_to_tmem)
  bias_to_tmem(tlane + kColD0 + unit0, s.brz0 + unit0);
  bias_to_tmem(tlane + kColD0 + 64 + unit0, s.brz0 + 64 + unit0);
  bias_to_tmem(tlane + kColD0 + 128 + unit0, s.bhn0 + unit0);
  bias_to_tmem(tlane + kColRz + unit0, s.brz1 + unit0);
  bias_to_tmem(tlane + kColRz + 64 + unit0, s.brz1 + 64 + unit0);
  bias_to_tmem(tlane + kColIn + unit0, s.bin1 + unit0);
  bias_to_tmem(tlane + kColHn + unit0, s.bhn1 + unit0);
  tmem_st_wait();

  uint32_t pa = 0, pb = 0;  // mbarrier phase parities
  float h0r[kTcUnits], h1r[kTcUnits];

  const long long n_tiles = (a.rows + kTcRows - 1) / kTcRows;

  // normalised action windows of a tile (w_nl.py:121): act[buf][row][j][u], j = 0 oldest
  auto load_windows = [&](long long tile_, int buf_) {
    const long long row0_ = tile_ * kTcRows;
    for (int i = tid; i < kTcRows * BG; i += kTcThreads) {
      const int r = i / BG, rem = i - r * BG, j = rem / gin, u = rem - j * gin;
      long long grow = row0_ + r;
      if (grow >= a.rows) grow = a.rows - 1;
      const long long k = grow / a.T;
      const int t = (int)(grow - k * a.T);
      const float v = u < a.hist_ch ? a.hist[((size_t)k * a.L + t + j) * a.hist_ch + u] : (float)(B - 1 - j);  // encode_obs_time channel
      s.act[buf_][r * 8 + rem] = (v - s.act_mean[u]) * s.act_inv_std[u];
    }
  };
  // A(0): layer 0 on the newest window entry (reversed order, w_nl.py:27) from the zero state - no MMA
  auto cell_a0 = [&](int buf_) {
    float x[GIN];
#pragma unroll
    for (int v = 0; v < GIN; ++v) x[v] = s.act[buf_][row * 8 + (B - 1) * gin + v];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int u0 = unit0 + 8 * c;
      float gr[8], gz[8], gn[8], bh[8];
      lds8(s.brz0 + u0, gr);
      lds8(s.brz0 + 64 + u0, gz);
      lds8(s.bin0 + u0, gn);
      lds8(s.bhn0 + u0, bh);
#pragma unroll
      for (int v = 0; v < GIN; ++v) {
        float w[8];
        lds8(s.w_ih0 + (0 * kMaxNu + v) * kTcHg + u0, w);
#pragma unroll
        for (int i = 0; i < 8; ++i) gr[i] = fmaf(w[i], x[v], gr[i]);
        lds8(s.w_ih0 + (1 * kMaxNu + v) * kTcHg + u0, w);
#pragma unroll
        for (int i = 0; i < 8; ++i) gz[i] = fmaf(w[i], x[v], gz[i]);
        lds8(s.w_ih0 + (2 * kMaxNu + v) * kTcHg + u0, w);
#pragma unroll
        for (int i = 0; i < 8; ++i) gn[i] = fmaf(w[i], x[v], gn[i]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) h0r[8 * c + i] = gru_unit<kSplit3>(gr[i], gz[i], gn[i], bh[i], 0.0f);
    }
  };
  // publish h0r as the A operand, then (one thread) start A(1): D0 += W_hh0 h0(0)
  auto publish_h0_and_issue_a1 = [&]() {
    store_operand<kSplit3>(s.h0[0], s.h0[1], row, unit0, h0r);
    fence_proxy_async_smem();
    tmem_st_wait();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      issue_gemm<kSplit3>(tmem + kColD0, a_h0_hi, a_h0_lo, w_hh0_hi, w_hh0_lo, 192, true);
      mma_commit(&s.bar_a);
    }
  };
  // B(0), input part only (h1 = 0): D1_rz += W_ih1[r,z] h0(0), D1_in += W_ih1[n] h0(0)
  auto issue_b0 = [&]() {
    if (tid == 0) {
      fence_after_sync();
      issue_gemm<kSplit3>(tmem + kColRz, a_h0_hi, a_h0_lo, w_ih1_hi, w_ih1_lo, 128, true);
      issue_gemm<kSplit3>(tmem + kColIn, a_h0_hi, a_h0_lo, w_ih1_hi + n_rows_off, w_ih1_lo + n_rows_off, 64, true);
      mma_commit(&s.bar_b);
    }
  };

  // Tiles are software-pipelined: the head of tile i+1 (window load, A(0), MMA A(1)) runs under the tail of tile i
  // (MMA B(B-1) and its epilogue), so no MMA latency is exposed between tiles.
  int buf = 0;
  if ((long long)blockIdx.x < n_tiles) {
    load_windows(blockIdx.x, 0);
    __syncthreads();
    cell_a0(0);
    publish_h0_and_issue_a1();
    issue_b0();
  }
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
    const long long row0 = tile * kTcRows;
    const long long next_tile = tile + gridDim.x;
    const bool has_next = next_tile < n_tiles;

    for (int st = 0; st < B; ++st) {
      if (st + 1 < B) {
        // ================= epilogue A(st+1)  ||  MMA B(st) =================
        mbar_wait(&s.bar_a, pa); pa ^= 1;
        fence_after_sync();
        float x[GIN];
#pragma unroll
        for (int v = 0; v < GIN; ++v) x[v] = s.act[buf][row * 8 + (B - 2 - st) * gin + v];
#pragma unroll
        for (int c = 0; c < 2; ++c) {  // 8 units at a time keeps the transient registers low
          const int u0 = unit0 + 8 * c;
          float gr[8], gz[8], ghn[8], gn[8];
          tmem_ld8(tlane + kColD0 + u0, gr);         // W_hr h + b_ir + b_hr
          tmem_ld8(tlane + kColD0 + 64 + u0, gz);    // W_hz h + b_iz + b_hz
          tmem_ld8(tlane + kColD0 + 128 + u0, ghn);  // W_hn h + b_hn
          lds8(s.bin0 + u0, gn);
          tmem_ld_wait();
#pragma unroll
          for (int v = 0; v < GIN; ++v) {
            float w[8];
            lds8(s.w_ih0 + (0 * kMaxNu + v) * kTcHg + u0, w);
#pragma unroll
            for (int i = 0; i < 8; ++i) gr[i] = fmaf(w[i], x[v], gr[i]);
            lds8(s.w_ih0 + (1 * kMaxNu + v) * kTcHg + u0, w);
#pragma unroll
            for (int i = 0; i < 8; ++i) gz[i] = fmaf(w[i], x[v], gz[i]);
            lds8(s.w_ih0 + (2 * kMaxNu + v) * kTcHg + u0, w);
#pragma unroll
            for (int i = 0; i < 8; ++i) gn[i] = fmaf(w[i], x[v], gn[i]);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) h0r[8 * c + i] = gru_unit<kSplit3>(gr[i], gz[i], gn[i], ghn[i], h0r[8 * c + i]);
          bias_to_tmem8(tlane + kColD0 + u0, s.brz0 + u0);
          bias_to_tmem8(tlane + kColD0 + 64 + u0, s.brz0 + 64 + u0);
          bias_to_tmem8(tlane + kColD0 + 128 + u0, s.bhn0 + u0);
        }
        // h0's image is still being read by MMA B(st): wait for its commit before overwriting
        mbar_wait(&s.bar_b, pb); pb ^= 1;
        fence_after_sync();
        store_operand<kSplit3>(s.h0[0], s.h0[1], row, unit0, h0r);
        fence_proxy_async_smem();
        tmem_st_wait();
        fence_before_sync();
        __syncthreads();
        if (tid == 0 && st + 2 < B) {  // A(st+2): D0 = W_hh0 h0(st+1)
          fence_after_sync();
          issue_gemm<kSplit3>(tmem + kColD0, a_h0_hi, a_h0_lo, w_hh0_hi, w_hh0_lo, 192, true);
          mma_commit(&s.bar_a);
        }
      } else {
        if (has_next) {  // head of the next tile, under MMA B(B-1)
          load_windows(next_tile, buf ^ 1);
          __syncthreads();
          cell_a0(buf ^ 1);
        }
        mbar_wait(&s.bar_b, pb); pb ^= 1;
        fence_after_sync();
        if (has_next) publish_h0_and_issue_a1();  // h0(B-1)'s readers (MMA B(B-1)) are done; D0 was reset after A(B-1)
      }
      // ================= epilogue B(st)  ||  MMA A(st+2) =================
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int u0 = unit0 + 8 * c;
        float sr[8], sz[8], gn[8], hn[8];
        tmem_ld8(tlane + kColRz + u0, sr);       // W_ir x + W_hr h + b_ir + b_hr
        tmem_ld8(tlane + kColRz + 64 + u0, sz);
        tmem_ld8(tlane + kColIn + u0, gn);       // W_in x + b_in
        tmem_ld8(tlane + kColHn + u0, hn);       // W_hn h + b_hn   (just b_hn at st == 0: no hidden product yet)
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) h1r[8 * c + i] = gru_unit<kSplit3>(sr[i], sz[i], gn[i], hn[i], st > 0 ? h1r[8 * c + i] : 0.0f);
        bias_to_tmem8(tlane + kColRz + u0, s.brz1 + u0);
        bias_to_tmem8(tlane + kColRz + 64 + u0, s.brz1 + 64 + u0);
        bias_to_tmem8(tlane + kColIn + u0, s.bin1 + u0);
        bias_to_tmem8(tlane + kColHn + u0, s.bhn1 + u0);
      }
      if (st + 1 < B) {
        store_operand<kSplit3>(s.h1[0], s.h1[1], row, unit0, h1r);
        fence_proxy_async_smem();
        tmem_st_wait();
        fence_before_sync();
        __syncthreads();
        if (tid == 0) {  // B(st+1): input part from h0(st+1), hidden part from h1(st)
          fence_after_sync();
          issue_gemm<kSplit3>(tmem + kColRz, a_h0_hi, a_h0_lo, w_ih1_hi, w_ih1_lo, 128, true);
          issue_gemm<kSplit3>(tmem + kColRz, a_h1_hi, a_h1_lo, w_hh1_hi, w_hh1_lo, 128, true);
          issue_gemm<kSplit3>(tmem + kColIn, a_h0_hi, a_h0_lo, w_ih1_hi + n_rows_off, w_ih1_lo + n_rows_off, 64, true);
          issue_gemm<kSplit3>(tmem + kColHn, a_h1_hi, a_h1_lo, w_hh1_hi + n_rows_off, w_hh1_lo + n_rows_off, 64, true);
          mma_commit(&s.bar_b);
        }
      }
    }
    // ================= linear_out on the top layer's last state (w_nl.py:29) =================
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int i = 0; i < kTcUnits; ++i) {
      o0 = fmaf(s.w_out[unit0 + i], h1r[i], o0);
      o1 = fmaf(s.w_out[kTcHg + unit0 + i], h1r[i], o1);
    }
    if (grp > 0) { s.pout[grp - 1][row * 2] = o0; s.pout[grp - 1][row * 2 + 1] = o1; }
    tmem_st_wait();
    fence_before_sync();  // this tile's TMEM loads / bias resets are ordered before the next tile's MMAs
    __syncthreads();
    if (grp == 0 && row0 + row < a.rows) {
      float2 o;
      o.x = ((o0 + s.pout[0][row * 2]) + s.pout[1][row * 2]) + s.pout[2][row * 2] + s.b_out[0];
      o.y = ((o1 + s.pout[0][row * 2 + 1]) + s.pout[1][row * 2 + 1]) + s.pout[2][row * 2 + 1] + s.b_out[1];
      *reinterpret_cast<float2*>(a.p_out + (row0 + row) * 2) = o;
    }
    if (has_next) issue_b0();  // D1 is free (read and reset above, ordered by the barrier)
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

int launch_encode_tc(nlc_model_s* m, const float* hist, int hist_ch, int K, int T, int B, float* p, int split3, cudaStream_t stream) {
  NLC_REQUIRE(B * m->gin <= 8, NLC_ERR_SHAPE, "tcgen05 encoder: window_length * input_width = %d exceeds 8", B * m->gin);
  NLC_REQUIRE(B >= 2, NLC_ERR_SHAPE, "tcgen05 encoder: window_length must be >= 2");
  EncTcArgs a;
  a.hist = hist; a.p_out = p; a.K = K; a.T = T; a.B = B; a.L = B - 1 + T; a.gin = m->gin; a.hist_ch = hist_ch;
  a.rows = (long long)K * T;
  a.m = m->d;
  const int smem = (int)sizeof(EncTcSmem) + 128;
  NLC_REQUIRE(m->gin == 1 || m->gin == 2, NLC_ERR_SHAPE, "tcgen05 encoder: GRU input width %d has no instantiation", m->gin);
  void (*kern)(EncTcArgs) = split3 ? (m->gin == 1 ? encode_tc_kernel<true, 1> : encode_tc_kernel<true, 2>)
                                   : (m->gin == 1 ? encode_tc_kernel<false, 1> : encode_tc_kernel<false, 2>);
  NLC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const long long n_tiles = (a.rows + kTcRows - 1) / kTcRows;
  const int grid = (int)(n_tiles < 148 ? n_tiles : 148);
  kern<<<grid, kTcThreads, smem, stream>>>(a);
  NLC_LAUNCH_OK("encode_tc_kernel");
  return NLC_OK;
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
