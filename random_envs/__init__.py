"""Drop-in module name: ``import random_envs`` resolves to the B200-native implementation.

The reference package is also called ``random_envs`` (its ``__init__`` imports the MuJoCo sub-package and the
cart-pole module, registering every env id as a side effect).  This alias does the same for the part of the suite
that is implemented here -- RandomCartPole-v0 and the DR samplers of every env id -- so the reference's README
snippet runs unchanged::

    import random_envs
    import gym                      # real gym if installed; otherwise call random_envs.install_gym() first
    env = gym.make('RandomCartPole-v0')

MuJoCo environments (RandomHopper-v0 ...) are out of scope: making them raises a KeyError / gym error; their xi
tables and samplers are available through ``random_envs.TaskSampler(env_id)``.
"""
from random_envs_b200 import *          # noqa: F401,F403
from random_envs_b200 import __all__ as _b200_all, __version__, gym, gym_compat  # noqa: F401
from . import random_cartpole, random_env  # noqa: F401


def install_gym():
    """Make ``import gym`` resolve to the bundled gym-0.21 subset when no real gym is installed."""
    return gym_compat.install_as_gym()


__all__ = list(_b200_all) + ["install_gym", "random_cartpole", "random_env"]
