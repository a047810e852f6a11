"""Neural Laplace dynamics model forward, restated from a reference ``state_dict`` (CPU oracle).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Follows ``w_nl.py`` of the reference:
* ``gru_history_encoder``  - ``ReverseGRUEncoder.forward`` ``w_nl.py:25-29`` (ctor ``:16-23``):
  flip the window in time, 2-layer GRU (hidden = hidden_units//2, ``w_nl.py:95``) from h=0,
  top-layer state after the last (= oldest) entry, then ``linear_out``.
  Gate order / equations are torch.nn.GRU's (r, z, n).
* ``laplace_rep``          - ``LaplaceRepresentationFunc.forward`` ``w_nl.py:55-63``.
* ``nl_forward``           - ``NeuralLaplaceModel.forward`` ``w_nl.py:117-145``.
The inverse Laplace transform is ``oracle.ilt`` (parity unpinned, see there).
"""
from __future__ import annotations

import math

import torch

from . import ilt

STATE_DICT_KEYS = (
    "state_mean", "state_std", "action_mean", "action_std", "dt",
    "action_encoder.gru.weight_ih_l0", "action_encoder.gru.weight_hh_l0",
    "action_encoder.gru.bias_ih_l0", "action_encoder.gru.bias_hh_l0",
    "action_encoder.gru.weight_ih_l1", "action_encoder.gru.weight_hh_l1",
    "action_encoder.gru.bias_ih_l1", "action_encoder.gru.bias_hh_l1",
    "action_encoder.linear_out.weight", "action_encoder.linear_out.bias",
    "laplace_rep_func.linear_tanh_stack.0.weight", "laplace_rep_func.linear_tanh_stack.0.bias",
    "laplace_rep_func.linear_tanh_stack.2.weight", "laplace_rep_func.linear_tanh_stack.2.bias",
    "laplace_rep_func.linear_tanh_stack.4.weight", "laplace_rep_func.linear_tanh_stack.4.bias",
)


def _gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    H = h.shape[-1]
    gi = x @ w_ih.T + b_ih
    gh = h @ w_hh.T + b_hh
    r = torch.sigmoid(gi[:, :H] + gh[:, :H])
    z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
    return (1.0 - z) * n + z * h


def gru_history_encoder(sd, window):
    """``window``: (K, B, nu[+1]) normalised actions, oldest first.  Returns (K, 2)."""
    pre = "action_encoder.gru."
    H = sd[pre + "weight_hh_l0"].shape[1]
    K, B, _ = window.shape
    h0 = window.new_zeros(K, H)
    h1 = window.new_zeros(K, H)
    for j in range(B - 1, -1, -1):  # reversed in time: newest entry first (w_nl.py:27)
        h0 = _gru_cell(window[:, j], h0, sd[pre + "weight_ih_l0"], sd[pre + "weight_hh_l0"],
                       sd[pre + "bias_ih_l0"], sd[pre + "bias_hh_l0"])
        h1 = _gru_cell(h0, h1, sd[pre + "weight_ih_l1"], sd[pre + "weight_hh_l1"],
                       sd[pre + "bias_ih_l1"], sd[pre + "bias_hh_l1"])
    return h1 @ sd["action_encoder.linear_out.weight"].T + sd["action_encoder.linear_out.bias"]


def laplace_rep(sd, inp, nx, S):
    """MLP on ``[theta_s | phi_s | p]`` -> (theta, phi), each (N, nx, S)  (w_nl.py:55-63)."""
    pre = "laplace_rep_func.linear_tanh_stack."
    x = inp.reshape(-1, inp.shape[-1])
    x = torch.tanh(x @ sd[pre + "0.weight"].T + sd[pre + "0.bias"])
    x = torch.tanh(x @ sd[pre + "2.weight"].T + sd[pre + "2.bias"])
    out = (x @ sd[pre + "4.weight"].T + sd[pre + "4.bias"]).view(-1, 2 * nx, S)
    theta = torch.tanh(out[:, :nx, :]) * math.pi
    # phi_scale = pi (w_nl.py:52-53); tanh*pi/2 - pi/2 + pi/2
    phi = torch.tanh(out[:, nx:, :]) * math.pi / 2.0 - math.pi / 2.0 + math.pi / 2.0
    return theta, phi


def nl_forward(sd, obs, act, ts, *, normalize=True, normalize_time=True, return_parts=False):
    """Predicted state difference (K, nx).  ``act``: (K, B, nu) env units, oldest first;
    ``ts``: (K, 1) seconds.  ``sd`` is a reference ``state_dict`` (tensors of obs's dtype)."""
    nx = obs.shape[-1]
    S = sd["laplace_rep_func.linear_tanh_stack.4.weight"].shape[0] // (2 * nx)
    if normalize:
        obs_n = (obs - sd["state_mean"]) / sd["state_std"]
        act_n = (act - sd["action_mean"]) / sd["action_std"]
        if normalize_time:
            ts = ts / (sd["dt"] * 8.0)
    else:
        obs_n = obs
        act_n = act / 3.0
    if act_n.dim() == 2:
        act_n = act_n.unsqueeze(1)
    p_action = gru_history_encoder(sd, act_n)
    p = torch.cat((obs_n, p_action), dim=1)
    out = ilt.laplace_reconstruct(lambda i: laplace_rep(sd, i, nx, S), p, ts, recon_dim=nx,
                                  ilt_algorithm="fourier", ilt_reconstruction_terms=S)
    out = torch.squeeze(out)
    if return_parts:
        return out, p_action
    return out


def cast_state_dict(sd, dtype):
    return {k: torch.as_tensor(v).to(dtype) for k, v in sd.items()}
