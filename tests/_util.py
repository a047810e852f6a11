"""Shared helpers for the parity tests: golden-vector loading and error metrics."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ENVS = ("oderl-pendulum", "oderl-cartpole", "oderl-acrobot")
S_TERMS = 17
DT = 0.05


def short(env):
    return env.split("-")[1]


def load(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def weights(env, calibrated=False, dtype=torch.float64, shift=-4.0):
    """Reference state_dict (golden, raw random init); ``calibrated`` applies the documented
    phi-bias shift of oracle/gen_golden.py."""
    from oracle.costs import ENV_DIMS

    sd = {k: torch.from_numpy(v.copy()) for k, v in load("weights_" + short(env)).items()}
    if calibrated:
        nx = ENV_DIMS[env][0]
        sd["laplace_rep_func.linear_tanh_stack.4.bias"][nx * S_TERMS:] += shift
    return {k: v.to(dtype) for k, v in sd.items()}


def relerr(ref, got):
    """max |ref-got| / max |ref|  (the tolerance form used for every floating-point bound)."""
    ref = torch.as_tensor(ref).detach().cpu().to(torch.float64)
    got = torch.as_tensor(got).detach().cpu().to(torch.float64)
    denom = max(ref.abs().max().item(), 1e-30)
    return (ref - got.reshape(ref.shape)).abs().max().item() / denom
