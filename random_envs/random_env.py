"""``random_envs.random_env`` of the reference -> the GPU-sampling base class (see random_envs_b200.random_env)."""
from random_envs_b200.random_env import RandomEnv, TaskSampler  # noqa: F401
