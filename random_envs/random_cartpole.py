"""``random_envs.random_cartpole`` of the reference -> the CUDA-backed class (see random_envs_b200.random_cartpole)."""
from random_envs_b200.random_cartpole import RandomCartPoleEnv  # noqa: F401
from random_envs_b200.vector_env import RandomCartPoleVecEnv  # noqa: F401
