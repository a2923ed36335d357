"""Minimal stand-in for gym 0.18 so the UNMODIFIED reference package imports in this container.

TEST INFRASTRUCTURE ONLY (used by oracle/gen_golden.py and oracle/replay_reference.py).
`gym` and `matplotlib` are not installed here and there is no network; the reference's hot path
uses only the handful of gym-0.18 behaviours reproduced below:

* ``gym.Env`` base class,
* ``gym.core.Wrapper``: copies ``action_space`` / ``observation_space`` / ``reward_range`` /
  ``metadata`` at construction and forwards every non-underscore attribute read to ``self.env``,
* ``gym.core.ObservationWrapper``: ``reset``/``step`` pipe the inner observation through
  ``self.observation(obs)``,
* ``gym.spaces.{Discrete, Box, Dict}`` as plain value holders,
* ``gym.envs.registration.{register, make}`` with ``entry_point='module:Class'`` and kwargs passed
  to the constructor; no TimeLimit wrapper (the reference registers no ``max_episode_steps``).
"""
from . import core, spaces, error, utils  # noqa: F401
from .core import Env, Wrapper, ObservationWrapper  # noqa: F401
from .envs.registration import make, register  # noqa: F401

__version__ = "0.18.0-stub"
