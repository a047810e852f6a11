// Host-side packing of weight matrices into the tcgen05 (UMMA) shared-memory operand image.
//
// Layout: K-major, no swizzle ("interleaved" canonical layout).  A [rows][K] fp16 operand is stored as
// 8x8 core matrices of 128 contiguous bytes (8 rows x 16 bytes); core matrices adjacent along K are
// kCoreBytes apart ("leading byte offset"), 8-row groups are (K/8)*kCoreBytes apart ("stride byte
// offset").  The same image is copied verbatim global -> shared by the kernels.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>
#include <stddef.h>

namespace nlc {

constexpr int kCoreBytes = 128;

// element offset (in halves) of (row, k) inside a [rows][K] K-major no-swizzle operand image
__host__ __device__ inline size_t tc_core_offset(int row, int k, int K) {
  return (size_t)(row >> 3) * (size_t)(K >> 3) * 64 + (size_t)(k >> 3) * 64 + (size_t)(row & 7) * 8 + (size_t)(k & 7);
}

// w: [rows][K] fp64 row-major (PyTorch Linear/GRU layout: rows = output features = the MMA's N).
// hi = fp16(w), lo = fp16(w - hi).
inline void tc_pack_weight_split(const double* w, int rows, int K, uint16_t* hi, uint16_t* lo) {
  for (int r = 0; r < rows; ++r)
    for (int k = 0; k < K; ++k) {
      double v = w[(size_t)r * K + k];
      __half h = __float2half_rn((float)v);
      double rem = v - (double)__half2float(h);
      __half l = __float2half_rn((float)rem);
      size_t off = tc_core_offset(r, k, K);
      hi[off] = *reinterpret_cast<uint16_t*>(&h);
      lo[off] = *reinterpret_cast<uint16_t*>(&l);
    }
}

}  // namespace nlc
