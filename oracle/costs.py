"""Running cost of the three delayed ODE-RL environments (CPU oracle).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

``running_cost = -(diff_obs_reward_(state, exp_reward=False) + diff_ac_reward_(action))``
as in the closure ``mppi_with_model.py:145-171``; reward formulas restated from
``envs/oderl/envs/ctpendulum.py:139-155``, ``ctcartpole.py:289-346`` (swing_up, with the
``state_constraint`` / ``change_goal`` options), ``ctacrobot.py:153-166,233-255`` and
``base_env.py:297-301`` (``trigonometric2angle``).  Constants: ``ac_rew_const`` /
``vel_rew_const`` = 0.01 / 0.01 (``base_env.py:28-29``) except acrobot 1e-4 / 1e-1
(``ctacrobot.py:110-111``); all link lengths 1.
"""
from __future__ import annotations

import torch

ENV_IDS = {"oderl-pendulum": 0, "oderl-cartpole": 1, "oderl-acrobot": 2}
ENV_DIMS = {"oderl-pendulum": (3, 1), "oderl-cartpole": (5, 1), "oderl-acrobot": (6, 2)}
ENV_ACT_HIGH = {"oderl-pendulum": 2.0, "oderl-cartpole": 3.0, "oderl-acrobot": 5.0}
# normalisation constants the models are built with (train_utils.py:189-200)
ENV_STATE_STD = {
    "oderl-pendulum": [0.70634571, 0.70784512, 2.89072771],
    "oderl-cartpole": [2.88646771, 11.54556671, 0.70729307, 0.70692035, 17.3199048],
    "oderl-acrobot": [0.70711024, 0.70710328, 0.7072186, 0.7069949, 2.88642115, 2.88627309],
}


def pendulum_cost(state, action):
    cos_th, sin_th, thdot = state[..., 0], state[..., 1], state[..., 2]
    state_reward = -(1.0 ** 2) * ((1 - cos_th) ** 2 + sin_th ** 2)
    velocity_reward = -(thdot ** 2)
    reward = state_reward + 0.01 * velocity_reward
    reward = reward + (-0.01 * torch.sum(action ** 2, -1))
    return -reward


def cartpole_cost(state, action, state_constraint=False, change_goal=False, change_goal_flipped=False):
    x, xdot, cos_th_len, sin_th_len, thetadot = (state[..., :1], state[..., 1:2], state[..., 2:3],
                                                 state[..., 3:4], state[..., 4:])
    length = 1.0
    ee_pos = torch.cat([x + sin_th_len, cos_th_len], -1)
    if change_goal:
        goal_x = 2.0 if change_goal_flipped else -2.0
    else:
        goal_x = 0.0
    err = ee_pos - torch.tensor([goal_x, length], dtype=torch.float32).to(state.device)
    if state_constraint:
        position_error = err[:, 0] ** 2 + torch.exp(err[:, 0] * 10.0 + 7.0)
        angle_error = err[:, 1] ** 2
        state_reward = -torch.sum(torch.cat((position_error.view(-1, 1), angle_error.view(-1, 1)), dim=1), -1)
    else:
        state_reward = -torch.sum(err ** 2, -1)
    velocity_reward = -torch.sum(xdot ** 2, -1) - torch.sum(thetadot ** 2, -1)
    reward = state_reward + 0.01 * velocity_reward
    reward = reward + (-0.01 * torch.sum(action ** 2, -1))
    return -reward


def _trig2angle(c, s):
    C = c ** 2 + s ** 2
    c, s = c / C, s / C
    return torch.atan2(s / C, c / C)


def acrobot_cost(state, action):
    th1 = _trig2angle(state[..., 0], state[..., 1])
    th2 = _trig2angle(state[..., 2], state[..., 3])
    vel1, vel2 = state[..., 4], state[..., 5]
    velocity_reward = -(vel1 ** 2) - vel2 ** 2
    p1x, p1y = -1.0 * torch.cos(th1), 1.0 * torch.sin(th1)
    p2x = p1x - 1.0 * torch.cos(th1 + th2)
    p2y = p1y + 1.0 * torch.sin(th1 + th2)
    state_reward = -((p2x - 1.0 - 1.0) ** 2) - p2y ** 2
    reward = state_reward + 1e-1 * velocity_reward
    reward = reward + (-1e-4 * torch.sum(action ** 2, -1))
    return -reward


def running_cost(env_name, **opts):
    if env_name == "oderl-pendulum":
        if opts:
            raise TypeError("pendulum reward takes no options (ctpendulum.py:139)")
        return pendulum_cost
    if env_name == "oderl-cartpole":
        return lambda s, a: cartpole_cost(s, a, **opts)
    if env_name == "oderl-acrobot":
        if opts:
            raise TypeError("acrobot reward takes no options (ctacrobot.py:233)")
        return acrobot_cost
    raise KeyError(env_name)
