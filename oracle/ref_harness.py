"""Import the reference's own planner / model classes from ``/root/reference`` (build container only).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  ``/root/reference`` does not exist on the
GPU box; nothing in the ``-m gpu`` tests, ``smoke()`` or ``bench.py`` imports this module.

* ``planners.mppi_delay.MPPIDelay`` imports as is.
* ``w_nl`` needs ``torchlaplace`` (absent, SURVEY finding 3): ``sys.modules['torchlaplace']`` is
  pre-seeded with a stub whose ``laplace_reconstruct`` is ``oracle.ilt.laplace_reconstruct`` - so
  the GRU encoder, the representation MLP and ``NeuralLaplaceModel.forward`` are the reference's
  code and only the ILT is the (unpinned) restatement.
* The env modules need ``gym``; the reward formulas come from ``oracle.costs``.
"""
from __future__ import annotations

import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("NLC_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "planners", "mppi_delay.py"))


def load():
    """Returns (MPPIDelay, w_nl module, oracle.py module of the reference)."""
    if not available():
        raise RuntimeError("reference tree not present")
    from . import ilt

    if "torchlaplace" not in sys.modules:
        stub = types.ModuleType("torchlaplace")
        stub.laplace_reconstruct = ilt.laplace_reconstruct
        sys.modules["torchlaplace"] = stub
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    import importlib.util

    def _load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE_ROOT, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    if "config" not in sys.modules:
        _load("config", "config.py")
    mppi_mod = sys.modules.get("_ref_mppi_delay") or _load("_ref_mppi_delay", "planners/mppi_delay.py")
    w_nl = sys.modules.get("_ref_w_nl") or _load("_ref_w_nl", "w_nl.py")
    ref_oracle = sys.modules.get("_ref_oracle") or _load("_ref_oracle", "oracle.py")
    return mppi_mod.MPPIDelay, w_nl, ref_oracle


def build_reference_model(env_name, seed=0, hidden_units=128, s_recon_terms=17, dt=0.05, out_scale=1.0, encode_obs_time=False):
    """Random-init reference ``NeuralLaplaceModel`` as ``train_utils.get_nl_model`` builds it
    (``train_utils.py:29-54,183-200``), ``.double()`` as ``mppi_with_model.py:101``.

    ``out_scale`` multiplies the last MLP layer (weight and bias) - used to generate a second,
    "small-delta" fixture family whose rollouts stay in the operating range of a trained model."""
    import numpy as np

    from .costs import ENV_ACT_HIGH, ENV_DIMS, ENV_STATE_STD

    _, w_nl, _ = load()
    nx, nu = ENV_DIMS[env_name]
    torch.manual_seed(seed)
    model = w_nl.NeuralLaplaceModel(
        nx, nu, nx, hidden_units=hidden_units, s_recon_terms=s_recon_terms, ilt_algorithm="fourier",
        encode_obs_time=encode_obs_time, state_mean=np.zeros(nx), state_std=np.array(ENV_STATE_STD[env_name]),
        action_mean=np.array([0] * nu), action_std=np.array([ENV_ACT_HIGH[env_name] / 2.0]),
        normalize=True, normalize_time=True, dt=dt,
    ).double()
    if out_scale != 1.0:
        with torch.no_grad():
            last = model.laplace_rep_func.linear_tanh_stack[4]
            last.weight.mul_(out_scale)
            last.bias.mul_(out_scale)
    return model
