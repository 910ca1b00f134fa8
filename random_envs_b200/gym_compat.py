"""The slice of gym==0.21.0 the RandomCartPole-v0 drop-in needs.

The reference plugs into gym's registry (``random_cartpole.py:291-296``) and users write
``gym.make('RandomCartPole-v0')`` (``README.md:52-66``).  gym is not installable here (no network),
so this module provides the same names with gym 0.21 semantics [restated from memory of gym 0.21]:

  Env, Wrapper (attribute forwarding), TimeLimit (elapsed >= max => info['TimeLimit.truncated'] =
  not done; done = True), spaces.Discrete / spaces.Box, utils.seeding.np_random, logger.warn,
  envs.register / register / make.

If a real ``gym`` is importable it is used instead: ``register`` forwards to it and ``make`` is
gym's own, so the env shows up in the real registry exactly like the reference's.

Usage mirroring the reference README::

    import random_envs_b200 as random_envs          # registers RandomCartPole-v0
    from random_envs_b200 import gym                # real gym if present, else this module
    env = gym.make('RandomCartPole-v0')
"""
import importlib
import sys
import warnings

import numpy as np

try:  # pragma: no cover - gym is absent from the build image
    import gym as _real_gym
    if getattr(_real_gym, "__oracle_shim__", False):
        _real_gym = None
except Exception:  # noqa: BLE001
    _real_gym = None

HAVE_REAL_GYM = _real_gym is not None


class Env:
    metadata = {"render.modes": []}
    reward_range = (-float("inf"), float("inf"))
    spec = None
    action_space = None
    observation_space = None

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def render(self, mode="human"):
        raise NotImplementedError

    def close(self):
        pass

    def seed(self, seed=None):
        return []

    @property
    def unwrapped(self):
        return self

    def __enter__(self):
        return self

    def __exit__(self, *args):
        self.close()
        return False


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self.action_space = env.action_space
        self.observation_space = env.observation_space
        self.reward_range = env.reward_range
        self.metadata = env.metadata

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError("attempted to get missing private attribute '{}'".format(name))
        return getattr(self.env, name)

    @property
    def spec(self):
        return self.env.spec

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def render(self, mode="human", **kwargs):
        return self.env.render(mode, **kwargs)

    def close(self):
        return self.env.close()

    def seed(self, seed=None):
        return self.env.seed(seed)


class TimeLimit(Wrapper):
    def __init__(self, env, max_episode_steps=None):
        super().__init__(env)
        if max_episode_steps is None and env.spec is not None:
            max_episode_steps = env.spec.max_episode_steps
        self._max_episode_steps = max_episode_steps
        self._elapsed_steps = None

    def step(self, action):
        assert self._elapsed_steps is not None, "Cannot call env.step() before calling reset()"
        observation, reward, done, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._elapsed_steps >= self._max_episode_steps:
            info["TimeLimit.truncated"] = not done
            done = True
        return observation, reward, done, info

    def reset(self, **kwargs):
        self._elapsed_steps = 0
        return self.env.reset(**kwargs)


class Discrete:
    def __init__(self, n):
        assert n >= 0
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)
        self.np_random = np.random.RandomState()

    def seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)
        return [seed]

    def sample(self):
        return int(self.np_random.randint(self.n))

    def contains(self, x):
        # gym 0.21: Python int (bool included, as an int subclass), or a 0-d numpy value of integer dtype
        if isinstance(x, int):
            as_int = x
        elif isinstance(x, (np.generic, np.ndarray)) and x.dtype.char in np.typecodes["AllInteger"] and x.shape == ():
            as_int = int(x)
        else:
            return False
        return 0 <= as_int < self.n

    __contains__ = contains

    def __repr__(self):
        return "Discrete(%d)" % self.n

    def __eq__(self, other):
        return isinstance(other, Discrete) and self.n == other.n


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        if shape is not None and np.isscalar(low):
            low = np.full(shape, low)
            high = np.full(shape, high)
        self.low = np.asarray(low, dtype=self.dtype)
        self.high = np.asarray(high, dtype=self.dtype)
        self.shape = self.low.shape
        self.np_random = np.random.RandomState()

    def seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)
        return [seed]

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

    def sample(self):
        finite = np.isfinite(self.low) & np.isfinite(self.high)
        out = self.np_random.normal(size=self.shape)
        out = np.where(finite, self.np_random.uniform(np.where(finite, self.low, 0), np.where(finite, self.high, 1)), out)
        return out.astype(self.dtype)

    def __repr__(self):
        return "Box(%s, %s, %s, %s)" % (self.low.min(), self.high.max(), self.shape, self.dtype)


class _Namespace:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def _np_random(seed=None):
    if seed is not None and not (isinstance(seed, (int, np.integer)) and 0 <= seed):
        raise ValueError("Seed must be a non-negative integer or omitted, not {}".format(seed))
    if seed is None:
        seed = int(np.random.SeedSequence().entropy % (2 ** 32))
    return np.random.RandomState(int(seed) % (2 ** 32)), int(seed)


def _warn(msg, *args):
    warnings.warn(("WARN: " + msg) % args if args else "WARN: " + msg, stacklevel=2)


class EnvSpec:
    def __init__(self, id, entry_point=None, max_episode_steps=None, reward_threshold=None, kwargs=None):
        self.id = id
        self.entry_point = entry_point
        self.max_episode_steps = max_episode_steps
        self.reward_threshold = reward_threshold
        self._kwargs = dict(kwargs or {})

    def make(self, **kwargs):
        if callable(self.entry_point):
            cls = self.entry_point
        else:
            mod_name, attr = self.entry_point.split(":")
            cls = getattr(importlib.import_module(mod_name), attr)
        kw = dict(self._kwargs)
        kw.update(kwargs)
        env = cls(**kw)
        env.unwrapped.spec = self
        if self.max_episode_steps is not None:
            env = TimeLimit(env, max_episode_steps=self.max_episode_steps)
        return env

    def __repr__(self):
        return "EnvSpec(%s)" % self.id


registry = {}


def register(id, **kwargs):
    """gym.envs.register: records the spec locally and in the real gym registry when there is one."""
    registry[id] = EnvSpec(id, **{k: kwargs[k] for k in ("entry_point", "max_episode_steps", "reward_threshold", "kwargs")
                                  if k in kwargs})
    if HAVE_REAL_GYM:  # pragma: no cover
        try:
            _real_gym.envs.register(id=id, **kwargs)
        except Exception:  # noqa: BLE001 - already registered
            pass


def make(id, **kwargs):
    if id not in registry:
        raise KeyError("No registered env with id: {}".format(id))
    return registry[id].make(**kwargs)


spaces = _Namespace(Discrete=Discrete, Box=Box)
logger = _Namespace(warn=_warn)
utils = _Namespace(seeding=_Namespace(np_random=_np_random))
envs = _Namespace(register=register, registry=registry)
wrappers = _Namespace(TimeLimit=TimeLimit)

if HAVE_REAL_GYM:  # pragma: no cover
    Env = _real_gym.Env
    spaces = _real_gym.spaces
    make = _real_gym.make


class _LazyVector:
    """``gym.vector``: resolved on first use (gym_vector imports this module)."""

    def __getattr__(self, name):
        from . import gym_vector
        return getattr(gym_vector, name)


vector = _LazyVector()


def install_as_gym():
    """Make ``import gym`` resolve to this module (only when no real gym exists)."""
    if not HAVE_REAL_GYM:
        sys.modules.setdefault("gym", sys.modules[__name__])
    return sys.modules["gym"]
