// Packed fp32x2 arithmetic of sm_100 (PTX add/mul/fma .f32x2 -> SASS FADD2/FMUL2/FFMA2): one issue slot for two
// fp32 lanes-per-thread.  The FMA pipe throughput is unchanged (tools/pipe_bench.cu: 120 thread-FMA/clk/SM either way);
// what halves is the number of warp-instructions, which is what the gate epilogues are short of next to the MUFU pipe.
#pragma once
#include <stdint.h>

namespace nlc {

typedef unsigned long long f2_t;  // two packed fp32 values (lo = element 0)

__device__ __forceinline__ f2_t pk2(float a, float b) { f2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ f2_t pk2u(uint32_t a, uint32_t b) { f2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ void upk2(f2_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f2_t fma2(f2_t a, f2_t b, f2_t c) { f2_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2_t add2(f2_t a, f2_t b) { f2_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2_t mul2(f2_t a, f2_t b) { f2_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

__device__ __forceinline__ float mufu_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// Reciprocal of two positive, finite, normal floats.
//   kIter == 0: two MUFU.RCP (1 ulp class).
//   kIter  > 0: on the FMA pipe - magic-constant seed (relative error <= 0.12) refined by kIter Newton steps
//               e = 1 - d x, x += x e :  0.12 -> 1.5e-2 -> 2.1e-4 -> 4.5e-8.  Three steps are fp32-class; two steps
//               (2e-4) are used only by the single-pass fp16 mode.  Frees the MUFU pipe, which bounds the gate epilogues.
template <int kIter>
__device__ __forceinline__ f2_t rcp2(f2_t d) {
  float d0, d1;
  upk2(d, d0, d1);
  if (kIter == 0) return pk2(mufu_rcp(d0), mufu_rcp(d1));
  f2_t x = pk2u(0x7EF311C7u - __float_as_uint(d0), 0x7EF311C7u - __float_as_uint(d1));
  const f2_t one = pk2(1.0f, 1.0f);
  const f2_t nd = pk2(-d0, -d1);
#pragma unroll
  for (int i = 0; i < kIter; ++i) {
    const f2_t e = fma2(nd, x, one);
    x = fma2(x, e, x);
  }
  return x;
}

}  // namespace nlc
