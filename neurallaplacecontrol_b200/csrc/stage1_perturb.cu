// Stage 1 of MPPIDelay.command: roll U, perturb, bound, effective noise, perturbation cost, and the
// env-unit action history the rollout consumes (planners/mppi_delay.py:199-200, 255-260, 319-344,
// 347-356).  One warp per sample, lanes over the horizon: every global access is a contiguous row.
//
// The elementwise chain is written with explicit round-to-nearest intrinsics (no FMA contraction) so
// that `perturbed` and `noise` are bit-identical to the same chain evaluated by PyTorch in fp32.
#include "common.cuh"

namespace nlc {

struct PerturbArgs {
  nlc_mppi_params p;
  const float* U_prev; float* U_cur; int roll;
  const float* noise_in; uint32_t seed_lo, seed_hi, call_lo, call_hi;
  const unsigned long long* call_ptr;  // when set, the call index is read from device memory (graph-captured planner steps)
  const float* action_buffer;
  float* perturbed; float* noise; float* hist; float* actions; float* pert_cost;
};

// Standard normals for (global sample, t): Philox4x32-10, counter = (idx_lo, idx_hi, call_lo, call_hi)
// with idx = global_k*T + t, key = seed; Box-Muller on (x0,x1) and (x2,x3).  oracle/philox.py is the
// numpy statement of the same generator.
__device__ __forceinline__ void philox_normals(uint64_t idx, const PerturbArgs& a, uint32_t call_lo, uint32_t call_hi, float z[4]) {
  uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), call_lo, call_hi};
  philox4x32_10(c, a.seed_lo, a.seed_hi);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float u1 = ((float)c[2 * h] + 0.5f) * 2.3283064365386963e-10f;  // (0, 1]
    float u2 = (float)c[2 * h + 1] * 2.3283064365386963e-10f;       // [0, 1]
    float r = sqrtf(-2.0f * logf(u1));
    float s, co;
    sincospif(2.0f * u2, &s, &co);
    z[2 * h] = r * co;
    z[2 * h + 1] = r * s;
  }
}

template <int NU>
__global__ void __launch_bounds__(256) perturb_kernel(PerturbArgs a) {
  const int T = a.p.T, B = a.p.B, L = B - 1 + T;
  const int lane = threadIdx.x & 31;
  const int warp_in_grid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  const float u_scale = a.p.u_scale;
  uint32_t call_lo = a.call_lo, call_hi = a.call_hi;
  if (a.call_ptr) {
    const unsigned long long cv = *a.call_ptr;
    call_lo = (uint32_t)cv; call_hi = (uint32_t)(cv >> 32);
  }

  if (a.roll >= 0 && blockIdx.x == 0) {  // publish the rolled U for stage 4 (mppi_delay.py:199-200)
    for (int i = threadIdx.x; i < T * NU; i += blockDim.x) {
      int t = i / NU, u = i - t * NU;
      a.U_cur[i] = a.roll ? ((t < T - 1) ? a.U_prev[i + NU] : a.p.u_init[u]) : a.U_prev[i];
    }
  }

  for (int k = warp_in_grid; k < a.p.K; k += n_warps) {
    const int64_t gk = a.p.k_offset + k;
    float pc = 0.0f;
    for (int t = lane; t < T; t += 32) {
      float Ut[NU], nz[NU];
#pragma unroll
      for (int u = 0; u < NU; ++u)
        Ut[u] = a.roll > 0 ? ((t < T - 1) ? a.U_prev[(t + 1) * NU + u] : a.p.u_init[u]) : a.U_prev[t * NU + u];
      const size_t e = ((size_t)k * T + t) * NU;
      if (a.noise_in) {
#pragma unroll
        for (int u = 0; u < NU; ++u) nz[u] = a.noise_in[e + u];
      } else {
        float z[4];
        philox_normals((uint64_t)gk * (uint64_t)T + (uint64_t)t, a, call_lo, call_hi, z);
#pragma unroll
        for (int u = 0; u < NU; ++u) {
          float acc = a.p.noise_mu[u];
#pragma unroll
          for (int v = 0; v <= u; ++v) acc = fmaf(a.p.sigma_chol[u * NU + v], z[v], acc);
          nz[u] = acc;
        }
      }
      float pert[NU], nb[NU];
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        float pa = __fadd_rn(Ut[u], nz[u]);                                   // :321
        if (a.p.sample_null_action && gk == a.p.k_total - 1) pa = 0.0f;      // :322-323
        float x = __fmul_rn(pa, u_scale);                                     // :325
        if (a.p.has_bounds) x = fmaxf(fminf(x, a.p.u_max[u]), a.p.u_min[u]);  // :347-353
        pert[u] = __fdiv_rn(x, u_scale);                                      // :326
        nb[u] = __fsub_rn(pert[u], Ut[u]);                                    // :328
        a.perturbed[e + u] = pert[u];
        a.noise[e + u] = nb[u];
        float h = __fmul_rn(u_scale, pert[u]);                                // :258 u_scale * perturbed
        a.hist[((size_t)k * L + (B - 1) + t) * NU + u] = h;
        if (a.actions) a.actions[e + u] = __fdiv_rn(h, u_scale);              // :340
      }
#pragma unroll
      for (int v = 0; v < NU; ++v) {  // action_cost = lambda * noise @ Sigma^-1   (:329-335)
        float ac = 0.0f;
#pragma unroll
        for (int u = 0; u < NU; ++u) {
          float n = a.p.noise_abs_cost ? fabsf(nb[u]) : nb[u];
          ac = fmaf(a.p.lambda_ * n, a.p.sigma_inv[u * NU + v], ac);
        }
        pc = fmaf(Ut[v], ac, pc);                                             // :343
      }
    }
    for (int j = lane; j < (B - 1) * NU; j += 32)  // :255-257 the known part of the history
      a.hist[(size_t)k * L * NU + j] = a.action_buffer[NU + j];
    pc = warp_sum(pc);
    if (lane == 0) a.pert_cost[k] = pc;
  }
}

// nlc_perturb with the call index optionally taken from device memory (call_index_dev != nullptr)
int perturb_launch(const nlc_mppi_params* p, const float* U_prev_dev, float* U_dev, int roll, const float* noise_in_dev,
                   uint64_t seed, uint64_t call_index, const unsigned long long* call_index_dev, const float* action_buffer_dev,
                   float* perturbed_dev, float* noise_dev, float* hist_dev, float* actions_dev, float* pert_cost_dev, void* stream);

}  // namespace nlc

using namespace nlc;

extern "C" int nlc_perturb(const nlc_mppi_params* p, const float* U_prev_dev, float* U_dev, int roll,
                           const float* noise_in_dev, uint64_t seed, uint64_t call_index,
                           const float* action_buffer_dev, float* perturbed_dev, float* noise_dev, float* hist_dev,
                           float* actions_dev, float* pert_cost_dev, void* stream) {
  return perturb_launch(p, U_prev_dev, U_dev, roll, noise_in_dev, seed, call_index, nullptr, action_buffer_dev, perturbed_dev,
                        noise_dev, hist_dev, actions_dev, pert_cost_dev, stream);
}

int nlc::perturb_launch(const nlc_mppi_params* p, const float* U_prev_dev, float* U_dev, int roll, const float* noise_in_dev,
                        uint64_t seed, uint64_t call_index, const unsigned long long* call_index_dev,
                        const float* action_buffer_dev, float* perturbed_dev, float* noise_dev, float* hist_dev,
                        float* actions_dev, float* pert_cost_dev, void* stream) {
  NLC_REQUIRE(p && U_prev_dev && action_buffer_dev && perturbed_dev && noise_dev && hist_dev && pert_cost_dev,
              NLC_ERR_ARG, "nlc_perturb: null pointer");
  NLC_REQUIRE(p->K >= 1 && p->T >= 1 && p->B >= 1, NLC_ERR_ARG, "nlc_perturb: K, T, B must be positive");
  NLC_REQUIRE(p->nu >= 1 && p->nu <= 4, NLC_ERR_SHAPE, "nlc_perturb: nu %d outside [1,4]", p->nu);
  NLC_REQUIRE(roll == 0 || U_dev != nullptr, NLC_ERR_ARG, "nlc_perturb: roll needs U_dev");
  PerturbArgs a;
  a.p = *p;
  a.U_prev = U_prev_dev; a.U_cur = U_dev; a.roll = U_dev ? roll : -1;
  a.noise_in = noise_in_dev;
  a.seed_lo = (uint32_t)seed; a.seed_hi = (uint32_t)(seed >> 32);
  a.call_lo = (uint32_t)call_index; a.call_hi = (uint32_t)(call_index >> 32);
  a.call_ptr = call_index_dev;
  a.action_buffer = action_buffer_dev;
  a.perturbed = perturbed_dev; a.noise = noise_dev; a.hist = hist_dev; a.actions = actions_dev; a.pert_cost = pert_cost_dev;
  const int warps_per_block = 8;
  int blocks = (p->K + warps_per_block - 1) / warps_per_block;
  if (blocks > 148 * 32) blocks = 148 * 32;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (p->nu) {
    case 1: perturb_kernel<1><<<blocks, 256, 0, s>>>(a); break;
    case 2: perturb_kernel<2><<<blocks, 256, 0, s>>>(a); break;
    case 3: perturb_kernel<3><<<blocks, 256, 0, s>>>(a); break;
    default: perturb_kernel<4><<<blocks, 256, 0, s>>>(a); break;
  }
  NLC_LAUNCH_OK("perturb_kernel");
  return NLC_OK;
}
