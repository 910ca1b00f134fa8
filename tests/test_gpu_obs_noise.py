"""the suite's "Noisy" variants -- obs = state + sqrt(noise_level) * N(0, I), drawn for every
observation (after a step and after a reset); the state itself is not perturbed
(jinja/random_hopper.py:28,107-108, random_walker2d.py:139-140, random_humanoid.py:193-204).

Checked here: (1) fp64 values follow the contract (Box-Muller of the oracle's Philox uniforms, purpose 4) to 1e-12;
(2) the true state of a noisy env is bit-identical to a noise-free env stepped with the same actions (noise never
feeds back); (3) the noise is N(0, noise_level) -- moments and KS -- and independent across steps and components;
(4) the drop-in id RandomCartPoleNoisy-v0 (sibling of RandomHopperNoisy-v0) and the host-buffer path.
"""
import math

import numpy as np
import pytest
import scipy.stats
import torch

import random_envs_b200 as random_envs
from random_envs_b200 import gym
from oracle import c_oracle, dr_port

pytestmark = pytest.mark.gpu

SEARCH = [2.0, 20.0, 0.5, 3.0, 0.05, 0.3, 0.1, 1.0]


def _np(t):
    return t.detach().cpu().numpy()


def _pair(n, dtype, noise_level, **kw):
    mk = lambda noisy: random_envs.RandomCartPoleVecEnv(n, dtype=dtype, seed=21, noisy=noisy, noise_level=noise_level, **kw)
    envs = mk(True), mk(False)
    for e in envs:
        e.set_dr_distribution("uniform", SEARCH); e.set_dr_training(True)
    return envs


def _expected_noise_f64(seed, env_id, tick, which):
    u = c_oracle.uniforms(seed, env_id, tick, c_oracle.PURPOSE_OBS, which, 4, np.float64)
    z = np.zeros(4)
    z[0], z[1] = dr_port.box_muller_f64(u[0] + 2.0 ** -53, u[1])
    z[2], z[3] = dr_port.box_muller_f64(u[2] + 2.0 ** -53, u[3])
    return z


@pytest.mark.parametrize("n", [1, 3, 1027])
def test_fp64_observation_follows_the_contract_and_state_is_untouched(n):
    level = 1e-4
    noisy, clean = _pair(n, "float64", level, max_episode_steps=9)
    o0 = _np(noisy.reset()).copy(); c0 = _np(clean.reset()).copy()
    assert np.array_equal(_np(noisy.state), c0)                       # reset state itself carries no noise
    for i in {0, n // 2, n - 1}:
        want = c0[i] + math.sqrt(level) * _expected_noise_f64(21, i, 0, 1)     # reset() ran at tick 0: draw "1"
        assert np.max(np.abs(o0[i] - want)) <= 1e-12
    for k in range(1, 25):
        a = clean.sample_actions().clone()
        on, rn, dn, infon = noisy.step(a)
        oc, rc, dc, infoc = clean.step(a)
        assert torch.equal(noisy.state, oc) and torch.equal(dn, dc) and torch.equal(rn, rc)
        assert torch.equal(infon["TimeLimit.truncated"], infoc["TimeLimit.truncated"])
        assert torch.equal(noisy.get_task(), clean.get_task()) and torch.equal(noisy.elapsed, clean.elapsed)
        on, oc, dn = _np(on), _np(oc), _np(dn)
        for i in {0, n // 2, n - 1}:
            which = 1 if dn[i] else 0                                   # a finished env shows its noisy RESET observation
            want = oc[i] + math.sqrt(level) * _expected_noise_f64(21, i, k, which)
            assert np.max(np.abs(on[i] - want)) <= 1e-12, (k, i)


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_noise_law_moments_ks_and_independence(dtype):
    n, level = 1 << 18, 1e-3
    noisy, clean = _pair(n, dtype, level)
    noisy.reset(); clean.reset()
    prev = None
    for k in range(3):
        a = clean.sample_actions().clone()
        on, _, _, _ = noisy.step(a)
        oc, _, _, _ = clean.step(a)
        z = (_np(on).astype(np.float64) - _np(oc).astype(np.float64)) / math.sqrt(level)
        assert abs(z.mean()) < 6 / math.sqrt(z.size) and abs(z.var() - 1.0) < 0.01
        for c in range(4):
            d = scipy.stats.kstest(z[:, c], "norm").statistic
            assert d < 1.95 / math.sqrt(n) + (2e-3 if dtype == "float32" else 0.0), (c, d)   # fp32: obs rounding
        corr = np.corrcoef(z.T)
        assert np.max(np.abs(corr - np.eye(4))) < 0.01
        if prev is not None:
            assert abs(np.mean(prev * z)) < 6 / math.sqrt(z.size)        # fresh draw every step
        prev = z


def test_noise_level_zero_and_attribute_update():
    noisy, clean = _pair(257, "float64", 0.0)
    assert torch.equal(noisy.reset(), clean.reset())
    a = clean.sample_actions().clone()
    assert torch.equal(noisy.step(a)[0], clean.step(a)[0])
    noisy.noise_level = 0.04                                             # a plain attribute in the reference
    a = clean.sample_actions().clone()
    d = _np(noisy.step(a)[0]) - _np(clean.step(a)[0])
    assert 0.15 < d.std() < 0.25


def test_noisy_dropin_id_and_host_buffer_path():
    env = gym.make("RandomCartPoleNoisy-v0")
    env.seed(3)
    assert env.unwrapped.noisy and env.unwrapped.noise_level == 1e-4
    obs = env.reset()
    state = np.array(env.unwrapped.state)
    assert obs.shape == (4,) and obs.dtype == np.float64 and np.all(np.abs(state) <= 0.05)
    assert 0 < np.max(np.abs(obs - state)) < 0.1
    obs2, r, done, info = env.step(1)
    state2 = np.array(env.unwrapped.state)
    assert r == 1.0 and not done and 0 < np.max(np.abs(obs2 - state2)) < 0.1
    quiet = gym.make("RandomCartPole-v0"); quiet.seed(3)
    assert not quiet.unwrapped.noisy and np.array_equal(quiet.reset(), state)   # same seed, same true trajectory
    assert np.array_equal(quiet.step(1)[0], state2)

    host, dev = (random_envs.RandomCartPoleVecEnv(515, dtype="float32", seed=4, noisy=True, noise_level=1e-2) for _ in range(2))
    host.reset(); dev.reset()
    for _ in range(5):
        a = dev.sample_actions().clone()
        o, r, d, _ = dev.step(a)
        ho, hr, hd, _ = host.step_host(_np(a))
        assert np.array_equal(ho, _np(o)) and np.array_equal(hd, _np(d))


@pytest.mark.parametrize("dtype,w", [("float32", (0.0, 0.0, 1.0, 0.0)), ("float64", (0.1, 0.1, 1.0, 0.3)),
                                     ("float64", (0.0, 0.0, 1.0, 0.0))])
@pytest.mark.parametrize("n,K,limit", [(1, 5, 500), (777, 60, 500), (4099, 33, 7)])
def test_noisy_fused_rollout_equals_noisy_steps_with_the_policy_on_the_observation(dtype, w, n, K, limit):
    """In the Noisy variant the in-kernel policy sees what step()/reset() returned (state + noise), not the state:
    K fused steps == K step() calls driven by a = [w . obs > 0], bit for bit, including the obs buffer left behind."""
    mk = lambda: random_envs.RandomCartPoleVecEnv(n, dtype=dtype, seed=31, noisy=True, noise_level=4e-4, max_episode_steps=limit)
    fused, stepped = mk(), mk()
    for e in (fused, stepped):
        e.set_dr_distribution("uniform", SEARCH); e.set_dr_training(True)
    obs = stepped.reset().clone(); fused.reset()
    for _ in range(3):                                        # start the rollout from a mix of episode ages
        a = stepped.sample_actions().clone()
        obs = stepped.step(a)[0]; fused.step(a)
    fused.rollout(w, 0.0, K)
    for _ in range(K):
        if dtype == "float32":
            act = (obs[:, 2] * w[2] > 0).to(torch.uint8)      # tie-free single-weight policy: torch == FMA chain
        else:
            acc = obs[:, 0] * w[0]
            for c in range(1, 4):
                acc = acc + obs[:, c] * w[c]
            act = (acc + 0.0 > 0).to(torch.uint8)
        obs = stepped.step(act)[0]
    assert torch.equal(fused.state, stepped.state) and torch.equal(fused.obs, stepped.obs)
    assert torch.equal(fused.get_task(), stepped.get_task()) and torch.equal(fused.elapsed, stepped.elapsed)
    assert torch.equal(fused.episode, stepped.episode)
    differs = (fused.obs - fused.state).abs().max()
    assert 0 < float(differs) < 0.2


def test_noisy_checkpoint_resume_is_exact():
    """state_dict carries the observation buffer: a resumed noisy env continues bit for bit (the first fused step of
    a rollout acts on the stored observation)."""
    mk = lambda seed: random_envs.RandomCartPoleVecEnv(1031, dtype="float32", seed=seed, noisy=True, noise_level=1e-3)
    env = mk(17)
    env.set_dr_distribution("uniform", SEARCH); env.set_dr_training(True)
    env.reset()
    for _ in range(5):
        env.step(env.sample_actions())
    sd = env.state_dict()
    twin = mk(999); twin.set_dr_distribution("uniform", SEARCH); twin.set_dr_training(True)
    twin.load_state_dict(sd)
    for e in (env, twin):
        e.rollout((0.0, 0.0, 1.0, 0.0), 0.0, 40)
        e.step(e.sample_actions())
    assert torch.equal(env.obs, twin.obs) and torch.equal(env.state, twin.state) and torch.equal(env.episode, twin.episode)
