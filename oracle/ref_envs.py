"""The reference's OWN reward functions, imported from ``/root/reference`` (build container only).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  ``/root/reference`` does not exist on the GPU box; nothing
in the ``-m gpu`` tests, ``smoke()`` or ``bench.py`` imports this module.

The env modules (``envs/oderl/envs/ct{pendulum,cartpole,acrobot}.py``) import ``gym``, ``torchdiffeq`` and
``TorchDiffEqPack`` at module level; none is installed here and none is touched by the reward code.  A meta-path
finder serves empty placeholder modules for exactly those names, so the env modules import, and the reward methods
``diff_obs_reward_`` / ``diff_ac_reward_`` are called on ``object.__new__(cls)`` instances that carry only the
attributes those methods read (``l`` / ``length`` / ``LINK_LENGTH_*`` / ``swing_up`` / ``vel_rew_const`` /
``ac_rew_const``, with the values the constructors set: ``base_env.py:28-29``, ``ctacrobot.py:110-111``,
``ctpendulum.py:60``, ``ctcartpole.py:80``).  ``running_cost(env)`` is then the closure of
``mppi_with_model.py:145-171`` over the reference's own arithmetic.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import sys
import types

from .ref_harness import REFERENCE_ROOT, available  # noqa: F401

_STUBBED = ("gym", "torchdiffeq", "TorchDiffEqPack", "pyglet", "imageio", "pyvirtualdisplay")


class _Placeholder(types.ModuleType):
    """Module whose every attribute is another placeholder (enough for ``from gym import spaces`` etc.)."""

    __path__: list = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _Placeholder(self.__name__ + "." + name)
        setattr(self, name, sub)
        return sub

    def __call__(self, *a, **k):
        raise RuntimeError(f"{self.__name__} is a placeholder (not installed); the reward code must not call it")


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUBBED:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _Placeholder(spec.name)

    def exec_module(self, module):
        if module.__name__ == "gym":
            module.Env = type("Env", (object,), {})  # base class of BaseEnv (base_env.py:13)


_ENV_CLASSES = None


def load_env_classes():
    """(CTPendulum, CTCartpole, CTAcrobot) classes of the reference."""
    global _ENV_CLASSES
    if _ENV_CLASSES is not None:
        return _ENV_CLASSES
    if not available():
        raise RuntimeError("reference tree not present")
    finder = _Finder()
    sys.meta_path.append(finder)  # last: only names that nothing else can import are served
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    try:
        pend = importlib.import_module("envs.oderl.envs.ctpendulum")
        cart = importlib.import_module("envs.oderl.envs.ctcartpole")
        acro = importlib.import_module("envs.oderl.envs.ctacrobot")
    finally:
        sys.meta_path.remove(finder)
    _ENV_CLASSES = (pend.CTPendulum, cart.CTCartpole, acro.CTAcrobot)
    return _ENV_CLASSES


def make_env(env_name):
    """Attribute-only instance of the reference env class (constructor NOT run: it needs gym)."""
    CTPendulum, CTCartpole, CTAcrobot = load_env_classes()
    if env_name == "oderl-pendulum":
        e = object.__new__(CTPendulum)
        e.l = 1.0                                   # ctpendulum.py:60
        e.vel_rew_const, e.ac_rew_const = 0.01, 0.01  # base_env.py:28-29
    elif env_name == "oderl-cartpole":
        e = object.__new__(CTCartpole)
        e.length = 1.0                              # ctcartpole.py:80
        e.swing_up = True
        e.vel_rew_const, e.ac_rew_const = 0.01, 0.01
    elif env_name == "oderl-acrobot":
        e = object.__new__(CTAcrobot)
        e.LINK_LENGTH_1 = e.LINK_LENGTH_2 = 1.0     # ctacrobot.py:57-58
        e.vel_rew_const, e.ac_rew_const = 1e-1, 1e-4  # ctacrobot.py:110-111
    else:
        raise KeyError(env_name)
    return e


def running_cost(env_name, state_constraint=False, change_goal=False, change_goal_flipped=False):
    """``mppi_with_model.py:145-171`` over the reference env's own reward methods."""
    e = make_env(env_name)
    if env_name != "oderl-cartpole" and (state_constraint or change_goal):
        raise TypeError("state_constraint / change_goal exist on the cartpole reward only (ctcartpole.py:289-297)")

    def cost(state, action):
        if state_constraint:
            r = e.diff_obs_reward_(state, exp_reward=False, state_constraint=True)
        elif change_goal:
            r = e.diff_obs_reward_(state, exp_reward=False, change_goal=True, change_goal_flipped=change_goal_flipped)
        else:
            r = e.diff_obs_reward_(state, exp_reward=False)
        return -(r + e.diff_ac_reward_(action))

    return cost
