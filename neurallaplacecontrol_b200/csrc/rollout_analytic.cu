// Stages 2+3 with the analytic delayed dynamics in the planner's `dynamics` slot
// (oracle.py:11-224 behind mppi_with_model.py:129-143): one thread per sample, the horizon in registers.
#include "common.cuh"
#include "env_cost.cuh"

namespace nlc {

template <int NX>
__global__ void __launch_bounds__(128) rollout_analytic_kernel(nlc_rollout_opts o, const float* __restrict__ state0,
                                                               int sps, const float* __restrict__ hist,
                                                               const float* __restrict__ pert_cost, int K, int T, int B,
                                                               int nu, float* __restrict__ cost, float* __restrict__ states) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const int L = B - 1 + T;
  float s[NX];
#pragma unroll
  for (int c = 0; c < NX; ++c) s[c] = sps ? state0[(size_t)k * NX + c] : state0[c];
  float acc = 0.0f;
  const float* h = hist + (size_t)k * L * nu;
  for (int t = 0; t < T; ++t) {
    // window = hist[t : t+B]; delayed action = window[-(delay+1)] (oracle.py:23,99,187)
    env_analytic_step(o, s, h + (size_t)(t + B - 1 - o.delay) * nu);
    acc += env_running_cost(o, s, h + (size_t)(t + B - 1) * nu, nu);  // mppi_delay.py:288-290
    if (states) {
#pragma unroll
      for (int c = 0; c < NX; ++c) states[((size_t)k * T + t) * NX + c] = s[c];
    }
  }
  cost[k] = acc + (pert_cost ? pert_cost[k] : 0.0f);
}

int launch_rollout_analytic(const nlc_rollout_opts* o, int nx, const float* state, int sps, const float* hist,
                            const float* pert_cost, int K, int T, int B, int nu, float* cost, float* states,
                            cudaStream_t stream) {
  const int grid = (K + 127) / 128;
  switch (nx) {
    case 3: rollout_analytic_kernel<3><<<grid, 128, 0, stream>>>(*o, state, sps, hist, pert_cost, K, T, B, nu, cost, states); break;
    case 5: rollout_analytic_kernel<5><<<grid, 128, 0, stream>>>(*o, state, sps, hist, pert_cost, K, T, B, nu, cost, states); break;
    case 6: rollout_analytic_kernel<6><<<grid, 128, 0, stream>>>(*o, state, sps, hist, pert_cost, K, T, B, nu, cost, states); break;
    default: set_error("analytic rollout: unsupported nx %d", nx); return NLC_ERR_SHAPE;
  }
  NLC_LAUNCH_OK("rollout_analytic_kernel");
  return NLC_OK;
}

}  // namespace nlc
