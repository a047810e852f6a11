"""The planner<->model glue of ``mppi_with_model.py`` as declarative handles.

The reference passes opaque Python closures to ``MPPIDelay`` (``dynamics`` ``mppi_with_model.py:103-143``,
``running_cost`` ``:145-171``).  A fused kernel cannot call Python, and this package has no CPU fallback, so the
B200 planner accepts exactly these recognisable handles and rejects anything else loudly.
"""
from __future__ import annotations

import torch

from . import _lib


class NLDynamics:
    """``state + model(state, window, ts_pred)`` with ``ts_pred = dt`` (``mppi_with_model.py:74,103-122``)."""

    kind = _lib.DYN_NEURAL_LAPLACE

    def __init__(self, model, dt=0.05):
        self.model = model
        self.dt = float(dt)
        self.delay = 0

    def __call__(self, state, window):
        ts = torch.full((state.shape[0], 1), self.dt, dtype=torch.float64, device=state.device)
        if self.model.encode_obs_time:  # mppi_with_model.py:110-119: window position B-1 .. 0 as an extra channel
            B = window.shape[1]
            tchan = torch.flip(torch.arange(B, device=window.device), (0,)).view(1, B, 1).to(window.dtype)
            window = torch.cat((window, tchan.repeat(window.shape[0], 1, 1)), dim=2)
        out = self.model(state, window, ts)
        return state + out.reshape(state.shape).to(state.dtype)


class AnalyticDelayDynamics:
    """The analytic delayed dynamics of ``oracle.py:11-224`` in the planner's dynamics slot
    (``mppi_with_model.py:129-143``): one Euler step driven by ``window[:, -(delay+1)]``."""

    kind = _lib.DYN_ANALYTIC_DELAY

    def __init__(self, env_name, delay, dt=0.05):
        if env_name not in _lib.ENV_IDS:
            raise KeyError(env_name)
        self.env_name = env_name
        self.delay = int(delay)
        self.dt = float(dt)
        self.model = None

    def __call__(self, state, window):
        raise RuntimeError("AnalyticDelayDynamics is evaluated inside the fused rollout kernel only")


class EnvRunningCost:
    """``-(env.diff_obs_reward_(state, exp_reward=False, ...) + env.diff_ac_reward_(action))``
    (``mppi_with_model.py:145-171``) for the three ODE-RL environments."""

    def __init__(self, env_name, state_constraint=False, change_goal=False, change_goal_flipped=False):
        if env_name not in _lib.ENV_IDS:
            raise KeyError(env_name)
        if env_name != "oderl-cartpole" and (state_constraint or change_goal):
            # ctpendulum.py:139 / ctacrobot.py:233 take only exp_reward; the reference would raise TypeError
            raise TypeError(f"{env_name} reward takes no state_constraint / change_goal options")
        self.env_name = env_name
        self.state_constraint = bool(state_constraint)
        self.goal_x = (2.0 if change_goal_flipped else -2.0) if change_goal else 0.0  # ctcartpole.py:312-319

    def __call__(self, state, action):
        raise RuntimeError("EnvRunningCost is evaluated inside the fused rollout kernel only")


def get_action(action_buffer, action, action_delay):
    """``mppi_with_model.py:25-28``: roll the buffer, append the new action, return the delayed one."""
    action_buffer = torch.roll(action_buffer, -1, dims=0)
    action_buffer[-1] = action
    return action_buffer, action_buffer[-(action_delay + 1)]


def noise_sigma_for(nu, sigma=1.0, dtype=torch.double):
    """Covariance the reference's callers build (``mppi_with_model.py:66-70``)."""
    gamma = sigma ** 2
    off = 0.5 * gamma
    return torch.ones((nu, nu), dtype=dtype) * off + torch.eye(nu, dtype=dtype) * (gamma - off)
