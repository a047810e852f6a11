"""Analytic delayed dynamics and the closed-loop environment step (CPU oracle).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Restates, in torch fp64 and observation form, the one-step explicit-Euler dynamics of the reference's ``oracle.py``:
``pendulum_dynamics_dt_delay`` (``oracle.py:177-224``), ``cartpole_dynamics_dt_delay`` (``:11-86``, ``friction=False``)
and ``acrobot_dynamics_dt_delay`` (``:89-174``), each driven by ``window[:, -(delay+1)]``; and ``step_env`` of
``mppi_with_model.py:193-216`` (``get_action`` + one such step + reward).  Pinned by ``tests/golden/env_step_*.npz`` and
``plan_oracledyn_*.npz``, which were generated with the reference's own functions (``oracle/gen_golden*.py``).
"""
from __future__ import annotations

import math

import torch

from . import costs, mppi


def _angle(c, s):
    """obs (cos, sin) -> angle exactly as the reference does it: two normalisations by C = c^2 + s^2, then atan2."""
    C = c * c + s * s
    c, s = c / C, s / C
    return torch.atan2(s / C, c / C), c, s


def analytic_step(env: str, state: torch.Tensor, window: torch.Tensor, dt: float, delay: int) -> torch.Tensor:
    """state (N, nx) observation form, window (N, B, nu) env units oldest first -> next state (N, nx)."""
    u = window[:, -(delay + 1), :]
    ts = dt
    if env == "oderl-pendulum":
        th, _, _ = _angle(state[:, 0], state[:, 1])
        thdot = state[:, 2]
        a = torch.clamp(u[:, 0], -2.0, 2.0)
        newth = th + thdot * ts
        newthdot = thdot + (-15.0 * torch.sin(th + math.pi) + 3.0 * a) * ts  # g=10, m=l=1: -3g/(2l) = -15, 3/(m l^2) = 3
        return torch.stack((torch.cos(newth), torch.sin(newth), newthdot), dim=1)
    if env == "oderl-cartpole":
        x, x_dot, theta_dot = state[:, 0], state[:, 1], state[:, 4]
        theta, cth, sth = _angle(state[:, 2], state[:, 3])
        gravity, force_mag, masspole, length = 9.8, 3.0, 0.1, 1.0
        total_mass, pml = masspole + 1.0, masspole * length
        force = torch.clamp(u[:, 0], -3.0, 3.0) * force_mag
        temp = (force + pml * theta_dot * theta_dot * sth) / total_mass
        thetaacc = (gravity * sth - cth * temp) / (length * (4.0 / 3.0 - masspole * cth * cth / total_mass))
        xacc = temp - pml * thetaacc * cth / total_mass
        new_theta = theta + theta_dot * ts
        return torch.stack((x + x_dot * ts, x_dot + xacc * ts, torch.cos(new_theta), torch.sin(new_theta), theta_dot + thetaacc * ts), dim=1)
    if env == "oderl-acrobot":
        th1, _, _ = _angle(state[:, 0], state[:, 1])
        th2, _, _ = _angle(state[:, 2], state[:, 3])
        d1_, d2_ = state[:, 4], state[:, 5]
        a = torch.clamp(u, -5.0, 5.0)
        m1 = m2 = l1 = I1 = I2 = 1.0
        lc1 = lc2 = 0.5
        g = 9.8
        d1 = m1 * lc1 ** 2 + m2 * (l1 ** 2 + lc2 ** 2 + 2 * l1 * lc2 * torch.cos(th2)) + I1 + I2
        d2 = m2 * (lc2 ** 2 + l1 * lc2 * torch.cos(th2)) + I2
        phi2 = m2 * lc2 * g * torch.cos(th1 + th2 - math.pi / 2.0)
        phi1 = (-m2 * l1 * lc2 * d2_ ** 2 * torch.sin(th2) - 2 * m2 * l1 * lc2 * d2_ * d1_ * torch.sin(th2)
                + (m1 * lc1 + m2 * l1) * g * torch.cos(th1 - math.pi / 2) + phi2)
        dd2 = (a[:, 0] + d2 / d1 * phi1 - m2 * l1 * lc2 * d1_ ** 2 * torch.sin(th2) - phi2) / (m2 * lc2 ** 2 + I2 - d2 ** 2 / d1)
        dd1 = -(a[:, 1] + d2 * dd2 + phi1) / d1
        n1, n2 = th1 + d1_ * ts, th2 + d2_ * ts
        return torch.stack((torch.cos(n1), torch.sin(n1), torch.cos(n2), torch.sin(n2), d1_ + dd1 * ts, d2_ + dd2 * ts), dim=1)
    raise KeyError(env)


def make_analytic_dynamics(env: str, dt: float, delay: int):
    """The ``dynamics`` closure of ``mppi_with_model.py:129-143`` over :func:`analytic_step`."""
    return lambda state, window: analytic_step(env, state, window, dt, delay)


def env_step(env: str, states: torch.Tensor, buffers: torch.Tensor, actions: torch.Tensor, delay: int, dt: float):
    """``step_env`` for I instances: returns (new states (I, nx), new buffers (I, B, nu), rewards (I))."""
    new_b = torch.roll(buffers, -1, dims=1)
    new_b[:, -1] = actions  # get_action, mppi_with_model.py:25-28
    applied = new_b[:, -(delay + 1)]
    new_s = analytic_step(env, states, new_b, dt, delay)
    return new_s, new_b, -costs.running_cost(env)(new_s, applied)


def closed_loop(env, plan_fn, states0, delay, n_steps, dt, B=4):
    """``loop()`` of ``mppi_with_model.py:244-317`` for I instances with ``plan_fn(i, state, buffer) -> action``."""
    I = states0.shape[0]
    nu = costs.ENV_DIMS[env][1]
    s, b = states0.clone(), torch.zeros(I, B, nu, dtype=states0.dtype)
    total = torch.zeros(I, dtype=states0.dtype)
    for _ in range(n_steps):
        a = torch.stack([plan_fn(i, s[i], b[i]) for i in range(I)])
        s, b, r = env_step(env, s, b, a, delay, dt)
        total += r
    return total, s


__all__ = ["analytic_step", "make_analytic_dynamics", "env_step", "closed_loop", "mppi"]
