"""The gym.vector-style numpy facade against the oracle's SyncVectorEnv port (gym 0.21 step_wait semantics):
same (obs, rewards, dones, infos) contract, TimeLimit(500) truncation in infos, auto-reset returning the reset obs."""
import numpy as np
import pytest

import random_envs_b200 as random_envs
from random_envs_b200 import gym
from oracle import cartpole_port as port

pytestmark = pytest.mark.gpu

SEARCH = [2.0, 20.0, 0.5, 3.0, 0.05, 0.3, 0.1, 1.0]


def test_surface_like_gym_vector_env():
    venv = gym.vector.make("RandomCartPole-v0", num_envs=7, dtype="float64")
    assert venv.num_envs == 7 and venv.observation_space.shape == (7, 4) and venv.single_action_space.n == 2
    assert venv.single_observation_space.shape == (4,) and venv.action_space.shape == (7,)
    venv.seed(3)
    venv.set_dr_distribution("uniform", SEARCH); venv.set_dr_training(True)
    obs = venv.reset()
    assert isinstance(obs, np.ndarray) and obs.shape == (7, 4) and np.all(np.abs(obs) <= 0.05)
    a = venv.action_space.sample()
    assert venv.action_space.contains(a)
    obs2, rew, done, infos = venv.step(a)
    assert obs2.shape == (7, 4) and rew.dtype == np.float64 and rew.shape == (7,) and done.dtype == np.bool_
    assert len(infos) == 7 and infos[0] == {} and list(infos) == [{}] * 7
    assert venv.get_task().shape == (7, 4) and np.all(venv.get_task() >= np.array(SEARCH[0::2]))
    with pytest.raises(AssertionError):
        venv.step(np.ones(7, dtype=np.float32))                  # Discrete(2) rejects floats (random_cartpole.py:173-174)
    with pytest.raises(AssertionError):
        venv.step(np.ones(6, dtype=np.int64))
    with pytest.raises(RuntimeError):
        venv.step_wait()
    with pytest.raises(KeyError):
        gym.vector.make("RandomHopper-v0", num_envs=2)
    assert gym.vector.make("RandomCartPoleNoisy-v0", num_envs=2).core.noisy
    venv.close()


def test_facade_matches_the_sync_vector_env_port_step_for_step():
    n, steps = 24, 700
    venv = random_envs.RandomCartPoleGymVectorEnv(n, dtype="float64", seed=5)
    venv.set_dr_distribution("uniform", SEARCH); venv.set_dr_training(True)
    obs = venv.reset()
    xi = venv.get_task()
    envs = []
    for i in range(n):
        e = port.TimeLimitPort(port.CartPolePort())
        e.reset()
        e.env.state = tuple(obs[i]); e.set_task(*xi[i])
        envs.append(e)
    w = np.array([0.1, 0.1, 1.0, 0.3])
    rs = np.random.RandomState(0)
    truncations = terminations = 0
    for k in range(steps):
        # half the envs follow a stabilising policy (they reach the 500-step limit), the others act randomly
        a = np.where(np.arange(n) % 2 == 0, (obs @ w > 0).astype(np.int64), rs.randint(0, 2, n))
        o_ref, r_ref, d_ref, t_ref = port.sync_vector_step(envs, [int(v) for v in a])
        obs, rew, done, infos = venv.step(a)
        assert np.array_equal(done, d_ref), k
        assert np.array_equal(rew, r_ref)
        assert [bool(infos[i].get("TimeLimit.truncated", False)) for i in range(n)] == list(t_ref)
        alive = ~done
        assert np.max(np.abs(obs[alive] - o_ref[alive]), initial=0.0) <= 1e-9, k
        assert np.all(np.abs(obs[done]) <= 0.05)                 # the returned obs of a finished env is its RESET obs
        xi = venv.get_task()
        for i in np.where(done)[0]:                              # different RNG streams: re-synchronise the port
            envs[i].env.state = tuple(obs[i]); envs[i].set_task(*xi[i])
        for i in np.where(alive)[0]:                             # and keep it teacher-forced (unstable mode e^{4t})
            envs[i].env.state = tuple(obs[i])
        truncations += int(t_ref.sum()); terminations += int((d_ref & ~t_ref).sum())
    assert truncations >= n // 2 and terminations > n             # both ends of TimeLimit were exercised
