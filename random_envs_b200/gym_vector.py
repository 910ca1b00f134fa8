"""``gym.vector``-style numpy facade over the device vector env.

The reference's batched CPU path is ``gym.vector.SyncVectorEnv([lambda: gym.make('RandomCartPole-v0')] * N)``
(gym 0.21; BASELINE.json north_star): numpy actions in; ``(obs (N, 4), rewards (N,) float64,
dones (N,) bool, infos)`` out; every sub-env under ``TimeLimit(500)``; a finished env is reset at once and its
returned observation is the reset observation.  That is exactly what ``renv_cartpole_step_*`` fuses, so this class
is only the numpy/host-buffer skin: ``step`` = pinned H2D of the actions -> one kernel -> D2H of obs/done
(``RandomCartPoleVecEnv.step_host``).  Extras a SyncVectorEnv does not have: the RandomEnv DR calls apply to all
sub-envs at once (``set_dr_distribution``, ``set_dr_training``, ``set_task``, ``get_task`` ...).

    from random_envs_b200 import gym
    venv = gym.vector.make('RandomCartPole-v0', num_envs=4096)
    venv.set_dr_distribution('uniform', [2, 20, 0.5, 3, 0.05, 0.3, 0.1, 1.0]); venv.set_dr_training(True)
    obs = venv.reset()
    obs, rewards, dones, infos = venv.step(venv.action_space.sample())
"""
import numpy as np

from . import gym_compat
from .vector_env import MAX_EPISODE_STEPS, THETA_THRESHOLD_RADIANS, X_THRESHOLD, RandomCartPoleVecEnv


class MultiDiscreteActions:
    """Batched ``Discrete(2)``: the action space gym.vector derives for N cart-poles."""

    def __init__(self, num_envs):
        self.nvec = np.full(num_envs, 2, dtype=np.int64)
        self.shape = (num_envs,)
        self.dtype = np.int64
        self._rng = np.random.RandomState()

    def seed(self, seed=None):
        self._rng = np.random.RandomState(seed)
        return [seed]

    def sample(self):
        return self._rng.randint(0, 2, size=self.shape).astype(np.int64)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and x.dtype.kind in "iub" and bool(((x == 0) | (x == 1)).all())


class InfoSequence:
    """``infos`` without building N dicts per step: ``infos[i]`` is ``{'TimeLimit.truncated': True}`` for an env that
    hit the step limit on this step and ``{}`` otherwise -- what gym 0.21's TimeLimit puts into the i-th info."""

    def __init__(self, truncated):
        self._truncated = truncated

    def __len__(self):
        return len(self._truncated)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        return {"TimeLimit.truncated": True} if self._truncated[i] else {}

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class RandomCartPoleGymVectorEnv:
    """``gym.vector.VectorEnv`` surface (gym 0.21) for N RandomCartPole-v0 envs on one GPU."""

    def __init__(self, num_envs, dtype="float32", noisy=False, **kwargs):
        self.num_envs = int(num_envs)
        self.core = RandomCartPoleVecEnv(num_envs, dtype=dtype, noisy=noisy, max_episode_steps=MAX_EPISODE_STEPS,
                                         auto_reset=True, track_truncated=True, **kwargs)
        high = np.array([X_THRESHOLD * 2, np.finfo(np.float32).max, THETA_THRESHOLD_RADIANS * 2,
                         np.finfo(np.float32).max], dtype=np.float32)
        self.single_observation_space = gym_compat.spaces.Box(-high, high, dtype=np.float32)
        self.single_action_space = gym_compat.spaces.Discrete(2)
        self.observation_space = gym_compat.spaces.Box(np.tile(-high, (self.num_envs, 1)), np.tile(high, (self.num_envs, 1)),
                                                       dtype=np.float32)
        self.action_space = MultiDiscreteActions(self.num_envs)
        self.closed = False
        self._pending = False

    # ---- gym.vector.VectorEnv ------------------------------------------------------------------------
    def seed(self, seeds=None):
        """One integer keys the whole batch (Philox: env i is stream (seed, i)); a list uses its first entry."""
        if isinstance(seeds, (list, tuple)):
            seeds = seeds[0] if seeds else None
        self.action_space.seed(seeds)
        return self.core.seed(seeds)

    def reset(self):
        self.reset_async()
        return self.reset_wait()

    def reset_async(self):
        self.core.reset()

    def reset_wait(self):
        return self.core.obs.cpu().numpy().copy()

    def step_async(self, actions):
        actions = np.asarray(actions)
        if actions.shape != (self.num_envs,) or actions.dtype.kind not in "iub":
            raise AssertionError("%r (%s) invalid" % (actions, type(actions)))      # Discrete(2).contains
        self.core.step_host_async(actions)
        self._pending = True

    def step_wait(self):
        if not self._pending:
            raise RuntimeError("Calling `step_wait` without any prior call to `step_async`.")   # gym's NoAsyncCallError text
        obs, reward, done, truncated = self.core.step_host_wait()
        self._pending = False
        # copies: the staging buffers are overwritten by the next step (gym returns fresh arrays too)
        return obs.copy(), reward.astype(np.float64), done.copy(), InfoSequence(truncated.copy())

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self, **kwargs):
        self.closed = True

    # ---- RandomEnv calls, applied to every sub-env -----------------------------------------------------
    def __getattr__(self, name):
        if name in ("set_dr_distribution", "get_dr_distribution", "set_dr_training", "get_dr_training", "set_random_task",
                    "sample_task", "sample_tasks", "load_dr_distribution_from_file", "get_task_search_bounds",
                    "set_task_search_bounds", "get_search_bounds_mean", "get_task_lower_bound", "dyn_index_to_name",
                    "get_reward_threshold", "task_dim", "min_task", "max_task", "mean_task", "stdev_task", "sampling",
                    "dr_training", "check_dr_violations"):
            return getattr(self.core, name)
        raise AttributeError(name)

    def get_task(self):
        """(N, 4) float64: one row per sub-env (``[env.get_task() for env in venv.envs]`` of a SyncVectorEnv)."""
        return self.core.get_task().double().cpu().numpy()

    def set_task(self, *task):
        self.core.set_task(*task)


def make(id, num_envs=1, asynchronous=True, wrappers=None, **kwargs):
    """``gym.vector.make``: only the ids whose dynamics live on the GPU can be batched here."""
    if wrappers is not None:
        raise NotImplementedError("wrappers are applied per Python env object; there are none here")
    if id == "RandomCartPole-v0":
        return RandomCartPoleGymVectorEnv(num_envs, **kwargs)
    if id == "RandomCartPoleNoisy-v0":
        return RandomCartPoleGymVectorEnv(num_envs, noisy=True, **kwargs)
    raise KeyError("No batched implementation for env id: {}".format(id))
