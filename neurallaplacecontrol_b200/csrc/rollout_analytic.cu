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
  for (int c = 0; c < NX; ++c) s[c] = state0[(size_t)(sps ? k / sps : 0) * NX + c];
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

// One closed-loop environment step for I independent instances (mppi_with_model.py:193-216, step_env): roll the action
// buffer and pick the delayed action (get_action, :25-28), advance the state by one explicit-Euler step of the true
// dynamics (base_env.py:136-173 integrate_system(2, ...) as stated analytically by oracle.py:11-224, observation form),
// reward of the new state with the applied action (base_env.py diff_reward = -running_cost).
template <int NX>
__global__ void env_step_kernel(nlc_rollout_opts o, float* __restrict__ state, float* __restrict__ abuf,
                                const float* __restrict__ action, int I, int B, int nu, float* __restrict__ reward) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= I) return;
  float* b = abuf + (size_t)i * B * nu;
  for (int j = 0; j + 1 < B; ++j)
    for (int u = 0; u < nu; ++u) b[j * nu + u] = b[(j + 1) * nu + u];
  for (int u = 0; u < nu; ++u) b[(B - 1) * nu + u] = action[(size_t)i * nu + u];
  const float* at = b + (size_t)(B - 1 - o.delay) * nu;
  float s[NX];
#pragma unroll
  for (int c = 0; c < NX; ++c) s[c] = state[(size_t)i * NX + c];
  env_analytic_step(o, s, at);
#pragma unroll
  for (int c = 0; c < NX; ++c) state[(size_t)i * NX + c] = s[c];
  if (reward) reward[i] = -env_running_cost(o, s, at, nu);
}

}  // namespace nlc

extern "C" int nlc_env_step(const nlc_rollout_opts* o, float* state_dev, float* action_buffer_dev, const float* action_dev,
                            int I, int B, int nu, float* reward_dev, void* stream) {
  using namespace nlc;
  NLC_REQUIRE(o && state_dev && action_buffer_dev && action_dev, NLC_ERR_ARG, "nlc_env_step: null pointer");
  NLC_REQUIRE(I >= 1 && B >= 1 && B <= 8, NLC_ERR_ARG, "nlc_env_step: I >= 1 and 1 <= B <= 8 required");
  NLC_REQUIRE(o->env >= NLC_ENV_PENDULUM && o->env <= NLC_ENV_ACROBOT, NLC_ERR_ARG, "unknown env id %d", o->env);
  NLC_REQUIRE(o->delay >= 0 && o->delay < B, NLC_ERR_ARG, "nlc_env_step: delay %d outside the %d-entry action buffer", o->delay, B);
  const int env_nu[3] = {1, 1, 2};
  NLC_REQUIRE(nu == env_nu[o->env], NLC_ERR_SHAPE, "env %d takes nu=%d actions, got %d", o->env, env_nu[o->env], nu);
  int dev = 0;
  NLC_CUDA_OK(cudaGetDevice(&dev));
  int rc = check_device_arch(dev);
  if (rc != NLC_OK) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = (I + 127) / 128;
  if (o->env == NLC_ENV_PENDULUM) env_step_kernel<3><<<grid, 128, 0, s>>>(*o, state_dev, action_buffer_dev, action_dev, I, B, nu, reward_dev);
  else if (o->env == NLC_ENV_CARTPOLE) env_step_kernel<5><<<grid, 128, 0, s>>>(*o, state_dev, action_buffer_dev, action_dev, I, B, nu, reward_dev);
  else env_step_kernel<6><<<grid, 128, 0, s>>>(*o, state_dev, action_buffer_dev, action_dev, I, B, nu, reward_dev);
  NLC_LAUNCH_OK("env_step_kernel");
  return NLC_OK;
}
