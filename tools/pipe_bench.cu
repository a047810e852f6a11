// Microbenchmark of the CUDA-core pipes the GRU / MLP gate epilogues live on (sm_100a): packed fp32x2 FMA issue rate,
// MUFU f16x2 forms, and complete GRU gate updates in the variants considered for encode_tc2.cu:
//   v0  5 MUFU per unit (2 ex2 + shared rcp for r,z; ex2 + rcp for n), scalar fp32           (round-1 kernel)
//   v1  same with log2(e) folded into the pre-activations and packed f32x2 arithmetic
//   v2  3 MUFU per unit: the two reciprocals as Newton iterations on the FMA pipe (packed)
//   v3  4 MUFU per unit: only the (r,z) reciprocal on the FMA pipe
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_bench pipe_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpf(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// Newton reciprocal of two positive floats on the FMA pipe: magic-constant seed (rel. error <= 12 %), 3 iterations
// e = 1 - d x ; x += x e   (12 % -> 1.5 % -> 2.2e-4 -> 5e-8)
__device__ __forceinline__ u64 rcp2_newton(u64 d) {
  float d0, d1;
  upk(d, d0, d1);
  const float x0 = __uint_as_float(0x7EF311C7u - __float_as_uint(d0)), x1 = __uint_as_float(0x7EF311C7u - __float_as_uint(d1));
  u64 x = pk(x0, x1);
  const u64 one = pk(1.0f, 1.0f);
  const u64 nd = d ^ 0x8000000080000000ull;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const u64 e = fma2(nd, x, one);
    x = fma2(x, e, x);
  }
  return x;
}

template <int V>
__device__ __forceinline__ void gru_pair(float pr0, float pr1, float pz0, float pz1, float gi0, float gi1, float gh0, float gh1,
                                         float& h0, float& h1) {
  if (V == 0) {
    float hh[2] = {h0, h1};
    const float pr[2] = {pr0, pr1}, pz[2] = {pz0, pz1}, gi[2] = {gi0, gi1}, gh[2] = {gh0, gh1};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float ea = ex2f(-1.44269504f * pr[i]), eb = ex2f(-1.44269504f * pz[i]);
      const float da = 1.0f + ea, db = 1.0f + eb;
      const float inv = rcpf(da * db);
      const float r = db * inv, z = da * inv;
      const float x = fmaf(r, gh[i], gi[i]);
      const float e = ex2f(-1.44269504f * (2.0f * x));
      const float n = fmaf(2.0f, rcpf(1.0f + e), -1.0f);
      hh[i] = fmaf(z, hh[i] - n, n);
    }
    h0 = hh[0]; h1 = hh[1];
  } else {
    // pre-activations arrive already multiplied by -log2(e) (r, z) and -2 log2(e) (n parts)
    const u64 one = pk(1.0f, 1.0f);
    const u64 ea = pk(ex2f(pr0), ex2f(pr1)), eb = pk(ex2f(pz0), ex2f(pz1));
    const u64 da = add2(ea, one), db = add2(eb, one);
    const u64 dd = mul2(da, db);
    u64 inv;
    if (V == 1) { float a, b; upk(dd, a, b); inv = pk(rcpf(a), rcpf(b)); } else inv = rcp2_newton(dd);
    const u64 r = mul2(db, inv), z = mul2(da, inv);
    const u64 x = fma2(r, pk(gh0, gh1), pk(gi0, gi1));
    float xa, xb;
    upk(x, xa, xb);
    const u64 dn = add2(pk(ex2f(xa), ex2f(xb)), one);
    u64 invn;
    if (V == 2) invn = rcp2_newton(dn); else { float a, b; upk(dn, a, b); invn = pk(rcpf(a), rcpf(b)); }
    const u64 n = fma2(pk(2.0f, 2.0f), invn, pk(-1.0f, -1.0f));
    const u64 hm = add2(pk(h0, h1), n ^ 0x8000000080000000ull);
    const u64 hn = fma2(z, hm, n);
    upk(hn, h0, h1);
  }
}

template <int V>
__global__ void __launch_bounds__(512, 1) gru_k(float* out, int iters, float seed) {
  float h[8], pre[8];
  for (int i = 0; i < 8; ++i) { h[i] = 0.1f * i + threadIdx.x * 1e-4f; pre[i] = seed * (i - 3.5f) + threadIdx.x * 1e-3f; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; i += 2)
      gru_pair<V>(pre[i] + h[i], pre[i + 1] + h[i + 1], pre[i] - h[i], pre[i + 1] - h[i + 1], pre[i] * 0.5f, pre[i + 1] * 0.5f,
                  h[i] + 0.25f, h[i + 1] + 0.25f, h[i], h[i + 1]);
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += h[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
__global__ void pipe_k(float* out, int iters, float seed) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-3f + i;
  u64 p[4];
  for (int i = 0; i < 4; ++i) p[i] = pk(a[2 * i], a[2 * i + 1]);
  const u64 c1 = pk(1.0001f, 0.9999f), c2 = pk(0.5f, 0.25f);
  uint32_t hx[8];
  for (int i = 0; i < 8; ++i) hx[i] = 0x3c003800u + threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = fmaf(a[i], 1.0001f, 0.5f);
      if (MODE == 1 && i < 4) { p[i] = fma2(p[i], c1, c2); p[i] = fma2(p[i], c1, c2); }  // 2 FFMA2 = 4 thread-FMAs per i<4 => 16 per inner = same FMAs as 2x MODE 0
      if (MODE == 2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(hx[i]));
      if (MODE == 3) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(hx[i]));
      if (MODE == 4) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); if (i < 4) { p[i] = fma2(p[i], c1, c2); p[i] = fma2(p[i], c1, c2); p[i] = fma2(p[i], c1, c2); p[i] = fma2(p[i], c1, c2); } }
      if (MODE == 5) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); a[i] = fmaf(a[i], 0.999f, 0.25f); a[i] = fmaf(a[i], 0.999f, 0.25f); a[i] = fmaf(a[i], 0.999f, 0.25f); a[i] = fmaf(a[i], 0.999f, 0.25f);
                       a[i] = fmaf(a[i], 0.999f, 0.25f); a[i] = fmaf(a[i], 0.999f, 0.25f); a[i] = fmaf(a[i], 0.999f, 0.25f); }
      if (MODE == 6) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(hx[i]));
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(hx[i]);
  for (int i = 0; i < 4; ++i) { float x, y; upk(p[i], x, y); s += x + y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static double run_kernel(void (*launch)(float*, int, int), int warps, int iters) {
  float* d;
  cudaMalloc(&d, 148 * 1024 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(d, 100, warps);
  cudaEventRecord(e0);
  launch(d, iters, warps);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaFree(d);
  return ms;
}
template <int MODE> void launch_pipe(float* d, int iters, int warps) { pipe_k<MODE><<<148, warps * 32>>>(d, iters, 1.0f); }
template <int V> void launch_gru(float* d, int iters, int warps) { gru_k<V><<<148, warps * 32>>>(d, iters, 0.3f); }

int main() {
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int iters = 20000;
  struct { const char* name; void (*fn)(float*, int, int); double ops; } pipes[] = {
      {"ffma (thread-FMA)", launch_pipe<0>, 8}, {"ffma2 (thread-FMA = 2/lane)", launch_pipe<1>, 16}, {"ex2.f16x2 (elements)", launch_pipe<2>, 16},
      {"tanh.f16x2 (elements)", launch_pipe<3>, 16}, {"tanh.bf16x2 (elements)", launch_pipe<6>, 16},
      {"1 ex2 + 2 ffma2 per thread-elt (count ex2)", launch_pipe<4>, 8}, {"1 ex2 + 7 ffma (count ex2)", launch_pipe<5>, 8}};
  for (int w : {4, 16, 32})
    for (auto& p : pipes) {
      const double ms = run_kernel(p.fn, w, iters);
      printf("%-44s warps=%2d %.3f ms  %.2f /clk/SM\n", p.name, w, ms, (double)iters * p.ops * w * 32 / (ms * 1e-3 * clk * 1e3));
    }
  struct { const char* name; void (*fn)(float*, int, int); } grus[] = {{"gru v0 (5 MUFU scalar)", launch_gru<0>}, {"gru v1 (5 MUFU packed, folded)", launch_gru<1>},
                                                                        {"gru v2 (3 MUFU + 2 Newton)", launch_gru<2>}, {"gru v3 (4 MUFU + 1 Newton)", launch_gru<3>}};
  for (int w : {8, 16})
    for (auto& g : grus) {
      const double ms = run_kernel(g.fn, w, iters);
      printf("%-44s warps=%2d %.3f ms  %.3f units/clk/SM  (%.1f clk per warp-unit per SMSP)\n", g.name, w, ms,
             (double)iters * 8 * w * 32 / (ms * 1e-3 * clk * 1e3), (ms * 1e-3 * clk * 1e3) / ((double)iters * 8 * w / 4));
    }
  return 0;
}
