"""Golden vectors for the closed-loop environment step: ``step_env`` of ``mppi_with_model.py:193-216``
- ``get_action`` (``:25-28``) + one explicit-Euler step of the true dynamics + reward - generated with the REFERENCE's
own ``oracle.py`` one-step dynamics (``oracle.py:11-224``; the env modules themselves need ``gym``, absent here) and the
restated rewards of ``oracle/costs.py``.

TEST INFRASTRUCTURE ONLY.  Run in the build container:  python -m oracle.gen_golden_envstep
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import costs, mppi, ref_harness
from .gen_golden import DT, ENVS, GOLDEN_DIR, START_STATE


def main():
    torch.set_grad_enabled(False)
    _, _, ref_oracle = ref_harness.load()
    for env in ENVS:
        nx, nu = costs.ENV_DIMS[env]
        short = env.split("-")[1]
        ah = costs.ENV_ACT_HIGH[env]
        fn = {"oderl-pendulum": ref_oracle.pendulum_dynamics_dt_delay, "oderl-cartpole": ref_oracle.cartpole_dynamics_dt_delay,
              "oderl-acrobot": ref_oracle.acrobot_dynamics_dt_delay}[env]
        cost = costs.running_cost(env)
        out = {}
        for delay in (0, 1, 3):
            g = torch.Generator().manual_seed(100 + delay)
            I, B, n_steps = 48, 4, 6
            state = torch.tensor(START_STATE[env], dtype=torch.float64).repeat(I, 1) + 0.05 * torch.randn(I, nx, generator=g, dtype=torch.float64)
            buf = (torch.rand(I, B, nu, generator=g, dtype=torch.float64) * 2 - 1) * ah
            actions = (torch.rand(n_steps, I, nu, generator=g, dtype=torch.float64) * 2 - 1) * ah
            ts = torch.full((I, 1), DT, dtype=torch.float64)
            states, rewards, bufs = [], [], []
            s, b = state.clone(), buf.clone()
            for it in range(n_steps):
                applied = torch.empty(I, nu, dtype=torch.float64)
                for i in range(I):  # get_action per instance (mppi_with_model.py:25-28)
                    b[i], applied[i] = mppi.get_action(b[i], actions[it, i], delay)
                s = fn(s, b, ts=ts, delay=delay, friction=False)  # window = the rolled buffer: picks b[:, -(delay+1)]
                r = -cost(s, applied)
                states.append(s.clone()); rewards.append(r.clone()); bufs.append(b.clone())
            out.update({f"d{delay}_state0": state, f"d{delay}_buf0": buf, f"d{delay}_actions": actions,
                        f"d{delay}_states": torch.stack(states), f"d{delay}_rewards": torch.stack(rewards),
                        f"d{delay}_bufs": torch.stack(bufs)})
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"env_step_{short}.npz"), **{k: v.numpy() for k, v in out.items()})
        print("wrote env_step_" + short)


if __name__ == "__main__":
    main()
