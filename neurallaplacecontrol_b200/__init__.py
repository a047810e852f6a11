"""B200-native MPPI planning hot path of NeuralLaplaceControl.

Mirrors the reference call surface for this path only:

* :class:`neurallaplacecontrol_b200.planners.mppi_delay.MPPIDelay`   <- ``planners/mppi_delay.py:52-381``
* :class:`neurallaplacecontrol_b200.w_nl.NeuralLaplaceModel`         <- ``w_nl.py:66-145``
* :mod:`neurallaplacecontrol_b200.closures` (``NLDynamics``, ``AnalyticDelayDynamics``, ``EnvRunningCost``,
  ``get_action``, ``noise_sigma_for``)                               <- ``mppi_with_model.py:25-28,66-70,103-171``
* :mod:`neurallaplacecontrol_b200.episode` (``BatchedMPPIDelay``, ``env_step``, ``run_closed_loop``): the closed loop
  of ``mppi_with_model.py:193-216,244-317`` for many instances at once (BASELINE config 5)

Everything computes in ``libnlc_b200.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/nlc_b200.h``).  There is no CPU fallback.
"""
from . import _lib  # noqa: F401
from .closures import AnalyticDelayDynamics, EnvRunningCost, NLDynamics, get_action, noise_sigma_for  # noqa: F401
from .episode import BatchedMPPIDelay, env_step, run_closed_loop  # noqa: F401
from .ilt import fourier_ilt  # noqa: F401
from .planners.mppi_delay import MPPIDelay  # noqa: F401
from .w_nl import NeuralLaplaceModel  # noqa: F401

__all__ = ["MPPIDelay", "NeuralLaplaceModel", "NLDynamics", "AnalyticDelayDynamics", "EnvRunningCost", "get_action",
           "noise_sigma_for", "fourier_ilt", "BatchedMPPIDelay", "env_step", "run_closed_loop"]
