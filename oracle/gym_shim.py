"""Minimal stand-in for gym==0.21.0, just enough to import the reference source.

TEST INFRASTRUCTURE (see oracle/__init__.py).  gym is pinned by the reference at
``requirements.txt:2`` but is not installed in this image and there is no
network.  The reference's hot-path files only touch a handful of gym names
(``random_cartpole.py:8-10,96-97,169,174,214,291-296``; ``random_env.py:1,4``):

    gym.Env, gym.spaces.Discrete, gym.spaces.Box, gym.logger.warn,
    gym.utils.seeding.np_random, gym.envs.register

Semantics restated from gym 0.21 from memory (the source is not in the container):
  * ``Discrete.contains``: Python ``int`` or 0-d integer-dtype numpy scalar in
    [0, n); anything else (floats, bools-as-np.bool_, arrays) is rejected.
  * ``seeding.np_random(seed)`` returns ``(RandomState, seed)``.  Real gym hashes
    the seed first; this shim seeds ``RandomState(seed)`` directly, so initial
    states are NOT comparable via seeds -- parity tests inject s0 explicitly
    (SURVEY.md section 8c "Seeding caveat").
  * ``envs.register`` records (id -> entry_point, max_episode_steps).
"""
import sys
import types

import numpy as np


class Env:
    metadata = {}
    reward_range = (-float("inf"), float("inf"))
    action_space = None
    observation_space = None


class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.np_random = np.random.RandomState()
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)
        return [seed]

    def sample(self):
        return int(self.np_random.randint(self.n))

    def contains(self, x):
        if isinstance(x, int):
            as_int = x
        elif isinstance(x, (np.generic, np.ndarray)) and (
            x.dtype.char in np.typecodes["AllInteger"] and x.shape == ()
        ):
            as_int = int(x)
        else:
            return False
        return 0 <= as_int < self.n


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.shape = self.low.shape
        self.dtype = np.dtype(dtype)


REGISTRY = {}


def _register(id, entry_point=None, max_episode_steps=None, kwargs=None, **_):
    REGISTRY[id] = dict(entry_point=entry_point, max_episode_steps=max_episode_steps,
                        kwargs=kwargs or {})


WARNINGS = []


def _warn(msg, *args):
    WARNINGS.append(msg % args if args else msg)


def _np_random(seed=None):
    return np.random.RandomState(seed), seed


def install():
    """Put the shim into ``sys.modules`` under the names the reference imports."""
    if "gym" in sys.modules and getattr(sys.modules["gym"], "__oracle_shim__", False):
        return sys.modules["gym"]
    gym = types.ModuleType("gym")
    gym.__oracle_shim__ = True
    gym.Env = Env

    spaces = types.ModuleType("gym.spaces")
    spaces.Discrete, spaces.Box = Discrete, Box
    logger = types.ModuleType("gym.logger")
    logger.warn = _warn
    utils = types.ModuleType("gym.utils")
    seeding = types.ModuleType("gym.utils.seeding")
    seeding.np_random = _np_random
    utils.seeding = seeding
    envs = types.ModuleType("gym.envs")
    envs.register = _register
    envs.registry = REGISTRY

    gym.spaces, gym.logger, gym.utils, gym.envs = spaces, logger, utils, envs
    for name, mod in [("gym", gym), ("gym.spaces", spaces), ("gym.logger", logger),
                      ("gym.utils", utils), ("gym.utils.seeding", seeding), ("gym.envs", envs)]:
        sys.modules[name] = mod
    return gym
