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


def action_relerr(ref_action, got_action, ref_U, u_scale):
    """|action - ref| relative to the magnitude of the planned control sequence it heads (max |U| * u_scale).

    The returned action U[0]*u_scale is a softmax-weighted mean of O(1) noise samples that largely cancel; with costs of
    magnitude ~50 at lambda = 1 one fp32 ulp of a cost (4e-6) moves its exponential weight by 4e-6, so an ABSOLUTE error
    of ~1e-5 on the action is the floor of any 32-bit evaluation (the CPU oracle run in fp32 is 100x further away).  A
    head that happens to be near zero would turn that floor into an arbitrarily large ratio to itself, so the action is
    measured against the scale of the sequence, like every other tensor (`relerr` divides by max |ref|)."""
    ref_action = torch.as_tensor(ref_action).detach().cpu().to(torch.float64).reshape(-1)
    got = torch.as_tensor(got_action).detach().cpu().to(torch.float64).reshape(-1)
    scale = float(torch.as_tensor(ref_U).detach().cpu().to(torch.float64).abs().max()) * float(u_scale)
    return (ref_action - got).abs().max().item() / max(scale, 1e-30)


def relerr_per_channel(ref, got):
    """max over the LAST axis' channels c of  max |ref_c - got_c| / max |ref_c|:  every state channel is held to the bound
    against its own magnitude (cos/sin channels of size 1 are not measured against velocities of size 10)."""
    ref = torch.as_tensor(ref).detach().cpu().to(torch.float64)
    got = torch.as_tensor(got).detach().cpu().to(torch.float64).reshape(ref.shape)
    C = ref.shape[-1]
    r, g = ref.reshape(-1, C), got.reshape(-1, C)
    denom = r.abs().amax(dim=0).clamp_min(1e-30)
    return float(((r - g).abs().amax(dim=0) / denom).max())


FULL_SIZE = {  # BASELINE configs 3 and 4 (SURVEY 8d inputs)
    "cfg3": ("oderl-cartpole", 8192, 30, "plan_cfg3_cartpole_K8192_H30"),
    "cfg4": ("oderl-acrobot", 65536, 50, "plan_cfg4_acrobot_K65536_H50"),
}
START_STATE = {
    "oderl-pendulum": [-1.0, 1.2246467991473532e-16, 1.0],
    "oderl-cartpole": [0.0, 0.0, -1.0, 1.2246467991473532e-16, 0.0],
    "oderl-acrobot": [1.0, 0.0, 1.0, 0.0, 0.0, 0.0],
}


def injected_noise(K, T, nu, seed=1, sigma=1.0):
    """The SURVEY 8d noise tensor: randn(K, T, nu; seed) @ chol(Sigma)^T in fp64 (same as oracle/gen_golden.py)."""
    from oracle import mppi

    g = torch.Generator().manual_seed(seed)
    z = torch.randn(K, T, nu, generator=g, dtype=torch.float64)
    return z @ torch.linalg.cholesky(mppi.noise_sigma_for(nu, sigma)).T
