"""Generate ``tests/golden/*.npz`` by running the REFERENCE's own code (build container only).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).   Usage:  ``python -m oracle.gen_golden``

What is the reference's own code in these vectors: ``MPPIDelay`` (noise injected by replacing
``noise_dist.sample``, the one sampling call per ``command()``, ``mppi_delay.py:319``),
``NeuralLaplaceModel`` with its GRU encoder and representation MLP, the ``dynamics`` closure shape
of ``mppi_with_model.py:103-122``, and ``oracle.py``'s analytic delayed dynamics.  What is NOT: the
inverse Laplace transform (``oracle.ilt`` restatement, parity unpinned).  The running cost is the
reference env classes' own ``diff_obs_reward_`` / ``diff_ac_reward_`` (``oracle/ref_envs.py`` imports
the env modules with placeholder ``gym`` / ``torchdiffeq`` modules); ``cost_pin.npz`` pins
``oracle.costs`` against them.

Weight families:
* ``raw``        - the reference modules' own random init under ``torch.manual_seed(0)``.
  Its rollouts explode (|delta state| ~ 1e2 per step, costs ~ 1e8): even the fp32 CPU oracle
  differs from fp64 by O(1) after 30 steps, so raw weights are used for single model steps and
  3-step rollouts only.
* ``calibrated`` - the same weights with ``PHI_BIAS_SHIFT`` added to the phi half of the last MLP
  layer's bias, which puts |F(s)| ~ 1e-2 and |delta state| ~ 0.1 per step (a trained model's
  operating range) so that whole-horizon plans are a meaningful parity target.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import costs, mppi, ref_envs, ref_harness

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
PHI_BIAS_SHIFT = -4.0
S_TERMS = 17
DT = 0.05
ENVS = ("oderl-pendulum", "oderl-cartpole", "oderl-acrobot")
START_STATE = {
    "oderl-pendulum": [-1.0, 1.2246467991473532e-16, 1.0],
    "oderl-cartpole": [0.0, 0.0, -1.0, 1.2246467991473532e-16, 0.0],
    "oderl-acrobot": [1.0, 0.0, 1.0, 0.0, 0.0, 0.0],
}


def calibrate_(sd, nx, S=S_TERMS, shift=PHI_BIAS_SHIFT):
    sd["laplace_rep_func.linear_tanh_stack.4.bias"][nx * S:] += shift
    return sd


def injected_noise(K, T, nu, seed=1, sigma=1.0):
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(K, T, nu, generator=g, dtype=torch.float64)
    return z @ torch.linalg.cholesky(mppi.noise_sigma_for(nu, sigma)).T


def _np(d):
    return {k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def reference_plan(model, env, K, T, U_init, state, action_buffer, noise, n_calls=1, dynamics=None):
    """Run the reference ``MPPIDelay.command`` ``n_calls`` times (buffer rolled with the returned
    action, delay 1) and return the planner's tensors after the last call."""
    MPPIDelay, _, _ = ref_harness.load()
    nx, nu = costs.ENV_DIMS[env]
    act_high = np.float32(costs.ENV_ACT_HIGH[env])  # gym Box bounds are float32 (mppi_with_model.py:57-58)
    ts_pred = torch.tensor(DT, dtype=torch.float64).view(1, 1).repeat(K, 1)

    if dynamics is None:
        def dynamics(state, window):
            return state + model(state, window, ts_pred)

    # the reward is the reference env class's own diff_obs_reward_ / diff_ac_reward_ (oracle/ref_envs.py)
    planner = MPPIDelay(dynamics, ref_envs.running_cost(env), nx, mppi.noise_sigma_for(nu), num_samples=K,
                        horizon=T, device="cpu", lambda_=1.0, u_min=torch.tensor(-act_high),
                        u_max=torch.tensor(act_high), u_scale=act_high, U_init=U_init.clone())
    out = None
    buf = action_buffer.clone()
    for call in range(n_calls):
        nz = noise[call] if noise.dim() == 4 else noise
        planner.noise_dist.sample = lambda shape, nz=nz: nz.clone()
        action = planner.command(np.asarray(state, dtype=np.float64), buf)
        out = {"noise": planner.noise, "perturbed_action": planner.perturbed_action,
               "cost_total": planner.cost_total, "cost_total_non_zero": planner.cost_total_non_zero,
               "omega": planner.omega, "states": planner.states, "actions": planner.actions,
               "U": planner.U, "action": action}
        buf, _ = mppi.get_action(buf, action, 1)
    return out


def cost_pin():
    """``cost_pin.npz``: the reference env classes' own reward arithmetic on seeded states/actions (every env, every
    cartpole option of ``mppi_with_model.py:145-171``) - pins ``oracle.costs`` (tests/test_oracle_golden.py)."""
    out = {}
    g = torch.Generator().manual_seed(40)
    for env in ENVS:
        nx, nu = costs.ENV_DIMS[env]
        short = env.split("-")[1]
        s = torch.randn(256, nx, generator=g, dtype=torch.float64) * torch.tensor(costs.ENV_STATE_STD[env])
        a = (torch.rand(256, nu, generator=g, dtype=torch.float64) * 2 - 1) * costs.ENV_ACT_HIGH[env]
        out[f"{short}_state"], out[f"{short}_action"] = s, a
        out[f"{short}_cost"] = ref_envs.running_cost(env)(s, a)
        if env == "oderl-cartpole":
            s2 = s.clone()
            s2[:, 0] = s2[:, 0] * 0.2 - 0.8  # keep exp(10 err + 7) of the state-constraint variant finite and varied
            out["cartpole_state_sc"] = s2
            out["cartpole_cost_state_constraint"] = ref_envs.running_cost(env, state_constraint=True)(s2, a)
            out["cartpole_cost_change_goal"] = ref_envs.running_cost(env, change_goal=True)(s, a)
            out["cartpole_cost_change_goal_flipped"] = ref_envs.running_cost(env, change_goal=True, change_goal_flipped=True)(s, a)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "cost_pin.npz"), **_np(out))


N_SPREAD = 64


def full_size_plan(env, K, T, family, name):
    """BASELINE configs 3 and 4 at their full K x H through the reference's own ``MPPIDelay`` + ``NeuralLaplaceModel``
    (fp64, SURVEY 8d inputs: zero U / buffer, start state, ``injected_noise(seed=1)``), outputs trimmed:
    every ``cost_total``, whole trajectories of ``N_SPREAD`` evenly spread samples, every final state (fp32 storage,
    2^-24 relative: far inside the 1e-4 bound), the 64 largest ``omega`` with their indices, ``U`` and the action.
    ``family``: ``cal`` = calibrated weights (trained-model operating range), ``raw`` = the reference modules' own
    random init, whose rollouts saturate (|state| ~ 1e4, costs ~ 1e7..1e9, omega one-hot) and are chaotic."""
    nx, nu = costs.ENV_DIMS[env]
    model = ref_harness.build_reference_model(env, seed=0, s_recon_terms=S_TERMS, dt=DT)
    if family == "cal":
        model.load_state_dict(calibrate_({k: v.clone() for k, v in model.state_dict().items()}, nx))
    out = reference_plan(model, env, K, T, torch.zeros(T, nu, dtype=torch.float64), START_STATE[env],
                         torch.zeros(4, nu, dtype=torch.float64), injected_noise(K, T, nu, seed=1))
    idx = torch.linspace(0, K - 1, N_SPREAD).round().long()
    top = torch.topk(out["omega"], 64)
    keep = {"spread_idx": idx, "states_spread": out["states"][idx], "cost_spread": out["cost_total"][idx],
            "U": out["U"], "action": out["action"], "argmin": int(out["cost_total"].argmin()),
            "noise_seed": 1, "phi_bias_shift": PHI_BIAS_SHIFT if family == "cal" else 0.0}
    if family == "cal":
        keep.update({"cost_total": out["cost_total"], "states_last": out["states"][:, -1].to(torch.float32),
                     "omega_top": top.values, "omega_top_idx": top.indices})
    # raw: the random-init dynamics are CHAOTIC (a 1e-16 perturbation grows ~6x per step: the fp64 oracle and the fp64
    # reference, which differ only in summation order, are 1e-13 apart after step 0 and O(1e3) apart after step 29), so
    # only the spread trajectories are kept; tests compare their first steps.
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **_np(keep))
    print(f"  {name}: cost [{float(out['cost_total'].min()):.4g}, {float(out['cost_total'].max()):.4g}]"
          f"  omega max {float(out['omega'].max()):.3g}  action {out['action'].tolist()}")


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_grad_enabled(False)
    torch.set_num_threads(4)
    _, _, ref_oracle = ref_harness.load()
    for env in ENVS:
        nx, nu = costs.ENV_DIMS[env]
        short = env.split("-")[1]
        model = ref_harness.build_reference_model(env, seed=0, s_recon_terms=S_TERMS, dt=DT)
        sd_raw = {k: v.detach().clone() for k, v in model.state_dict().items()}
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"weights_{short}.npz"), **_np(sd_raw))

        # --- single model steps, raw weights, reference NeuralLaplaceModel.forward -------------
        g = torch.Generator().manual_seed(10)
        K = 80
        obs = torch.randn(K, nx, generator=g, dtype=torch.float64) * torch.tensor(costs.ENV_STATE_STD[env])
        act = (torch.rand(K, 4, nu, generator=g, dtype=torch.float64) * 2 - 1) * costs.ENV_ACT_HIGH[env]
        ts_fixed = torch.full((K, 1), DT, dtype=torch.float64)
        ts_irreg = 0.01 + 0.39 * torch.rand(K, 1, generator=g, dtype=torch.float64)
        enc = model.action_encoder((act - model.action_mean) / model.action_std)
        fwd = {"obs": obs, "act": act, "ts_fixed": ts_fixed, "ts_irreg": ts_irreg,
               "p_action": enc, "out_fixed": model(obs, act, ts_fixed), "out_irreg": model(obs, act, ts_irreg)}
        # representation MLP alone (reference LaplaceRepresentationFunc.forward)
        rep_in = torch.randn(K, 2 * S_TERMS + nx + 2, generator=g, dtype=torch.float64)
        theta, phi = model.laplace_rep_func(rep_in)
        fwd.update({"rep_in": rep_in, "rep_theta": theta, "rep_phi": phi})
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"model_fwd_{short}.npz"), **_np(fwd))

        # --- raw weights, 3-step plan -----------------------------------------------------------
        K, T = 64, 3
        noise = injected_noise(K, T, nu, seed=3)
        U0 = torch.zeros(T, nu, dtype=torch.float64)
        buf = torch.zeros(4, nu, dtype=torch.float64)
        out = reference_plan(model, env, K, T, U0, START_STATE[env], buf, noise)
        out.update({"in_noise": noise, "in_U": U0, "in_state": np.array(START_STATE[env]), "in_buffer": buf})
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"plan_raw_{short}.npz"), **_np(out))

        # --- calibrated weights -----------------------------------------------------------------
        sd_cal = calibrate_({k: v.clone() for k, v in sd_raw.items()}, nx)
        model.load_state_dict(sd_cal)

        # ragged K, non-zero U / buffer / generic state, two consecutive control steps
        K, T = 100, 7
        g = torch.Generator().manual_seed(20)
        noise = torch.stack([injected_noise(K, T, nu, seed=4), injected_noise(K, T, nu, seed=5)])
        U0 = torch.randn(T, nu, generator=g, dtype=torch.float64) * 0.3
        buf = (torch.rand(4, nu, generator=g, dtype=torch.float64) * 2 - 1) * costs.ENV_ACT_HIGH[env]
        st = np.array(START_STATE[env]) + 0.05 * torch.randn(nx, generator=g, dtype=torch.float64).numpy()
        for n_calls in (1, 2):
            out = reference_plan(model, env, K, T, U0, st, buf, noise, n_calls=n_calls)
            out.update({"in_noise": noise, "in_U": U0, "in_state": st, "in_buffer": buf,
                        "phi_bias_shift": PHI_BIAS_SHIFT})
            np.savez_compressed(os.path.join(GOLDEN_DIR, f"plan_cal_{short}_calls{n_calls}.npz"), **_np(out))

        # per-sample start states (K, nx) (mppi_delay.py:243-244)
        stK = torch.tensor(st).view(1, -1) + 0.05 * torch.randn(K, nx, generator=g, dtype=torch.float64)
        out = reference_plan(model, env, K, T, U0, stK.numpy(), buf, noise[0])
        out.update({"in_noise": noise[0], "in_U": U0, "in_state": stK, "in_buffer": buf,
                    "phi_bias_shift": PHI_BIAS_SHIFT})
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"plan_cal_{short}_stateK.npz"), **_np(out))

        # --- reference oracle.py dynamics in the same planner slot ----------------
        fn = {"oderl-pendulum": ref_oracle.pendulum_dynamics_dt_delay,
              "oderl-cartpole": ref_oracle.cartpole_dynamics_dt_delay,
              "oderl-acrobot": ref_oracle.acrobot_dynamics_dt_delay}[env]
        for delay in (0, 1, 3):
            K, T = 100, 12
            ts_pred = torch.full((K, 1), DT, dtype=torch.float64)
            dyn = (lambda s, w, fn=fn, delay=delay, ts_pred=ts_pred: fn(s, w, ts=ts_pred, delay=delay, friction=False))
            noise = injected_noise(K, T, nu, seed=6 + delay)
            out = reference_plan(None, env, K, T, U0.new_zeros(T, nu), START_STATE[env], buf, noise, dynamics=dyn)
            out.update({"in_noise": noise, "in_U": U0.new_zeros(T, nu), "in_state": np.array(START_STATE[env]),
                        "in_buffer": buf, "delay": delay})
            np.savez_compressed(os.path.join(GOLDEN_DIR, f"plan_oracledyn_{short}_d{delay}.npz"), **_np(out))

    # --- encode_obs_time=True models: the closure of mppi_with_model.py:110-119 appends the window position
    #     B-1..0 as an extra GRU input channel; only nu = 1 envs broadcast through the reference's normalisation ----------
    for env in ("oderl-pendulum", "oderl-cartpole"):
        nx, nu = costs.ENV_DIMS[env]
        short = env.split("-")[1]
        model = ref_harness.build_reference_model(env, seed=1, s_recon_terms=S_TERMS, dt=DT, encode_obs_time=True)
        sd = calibrate_({k: v.detach().clone() for k, v in model.state_dict().items()}, nx)
        model.load_state_dict(sd)
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"weights_eot_{short}.npz"), **_np(sd))
        g = torch.Generator().manual_seed(30)
        K, T, B = 100, 6, 4
        tchan = torch.flip(torch.arange(B), (0,)).view(1, B, 1).to(torch.float64)
        obs = torch.randn(K, nx, generator=g, dtype=torch.float64) * torch.tensor(costs.ENV_STATE_STD[env])
        act = (torch.rand(K, B, nu, generator=g, dtype=torch.float64) * 2 - 1) * costs.ENV_ACT_HIGH[env]
        act_t = torch.cat((act, tchan.repeat(K, 1, 1)), dim=2)
        ts_fixed = torch.full((K, 1), DT, dtype=torch.float64)
        out = {"obs": obs, "act": act_t, "out_fixed": model(obs, act_t, ts_fixed),
               "p_action": model.action_encoder((act_t - model.action_mean) / model.action_std)}
        noise = injected_noise(K, T, nu, seed=31)
        U0 = torch.randn(T, nu, generator=g, dtype=torch.float64) * 0.3
        buf = (torch.rand(B, nu, generator=g, dtype=torch.float64) * 2 - 1) * costs.ENV_ACT_HIGH[env]
        ts_pred = torch.full((K, 1), DT, dtype=torch.float64)

        def dyn(state, window, model=model, ts_pred=ts_pred, tchan=tchan):
            w = torch.cat((window, tchan.repeat(window.shape[0], 1, 1)), dim=2)
            return state + model(state, w, ts_pred)

        plan = reference_plan(None, env, K, T, U0, START_STATE[env], buf, noise, dynamics=dyn)
        out.update({"plan_" + k: v for k, v in plan.items()})
        out.update({"in_noise": noise, "in_U": U0, "in_state": np.array(START_STATE[env]), "in_buffer": buf})
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"eot_{short}.npz"), **_np(out))

    # --- BASELINE config 1: pendulum K=1000 H=20 (calibrated weights), trimmed outputs ----------
    env = "oderl-pendulum"
    nx, nu = costs.ENV_DIMS[env]
    model = ref_harness.build_reference_model(env, seed=0, s_recon_terms=S_TERMS, dt=DT)
    model.load_state_dict(calibrate_({k: v.clone() for k, v in model.state_dict().items()}, nx))
    K, T = 1000, 20
    out = reference_plan(model, env, K, T, torch.zeros(T, nu, dtype=torch.float64), START_STATE[env],
                         torch.zeros(4, nu, dtype=torch.float64), injected_noise(K, T, nu, seed=1))
    keep = {"cost_total": out["cost_total"], "omega": out["omega"], "U": out["U"], "action": out["action"],
            "states_first16": out["states"][:16], "states_last": out["states"][:, -1],
            "noise_seed": 1, "phi_bias_shift": PHI_BIAS_SHIFT}
    np.savez_compressed(os.path.join(GOLDEN_DIR, "plan_cfg1_pendulum_K1000_H20.npz"), **_np(keep))
    cost_pin()
    for env, K, T, name in (("oderl-cartpole", 8192, 30, "cfg3_cartpole_K8192_H30"),
                            ("oderl-acrobot", 65536, 50, "cfg4_acrobot_K65536_H50")):
        for family in ("cal", "raw"):
            full_size_plan(env, K, T, family, f"plan_{name}_{family}")
    print("golden vectors written to", GOLDEN_DIR)
    for f in sorted(os.listdir(GOLDEN_DIR)):
        print(f"  {f}  {os.path.getsize(os.path.join(GOLDEN_DIR, f)) / 1024:.0f} KiB")


def main_full_size_only():
    torch.set_grad_enabled(False)
    for env, K, T, name in (("oderl-cartpole", 8192, 30, "cfg3_cartpole_K8192_H30"),
                            ("oderl-acrobot", 65536, 50, "cfg4_acrobot_K65536_H50")):
        for family in ("cal", "raw"):
            full_size_plan(env, K, T, family, f"plan_{name}_{family}")


if __name__ == "__main__":
    import sys

    main_full_size_only() if "--full-size-only" in sys.argv else main()
