"""Parity at BASELINE.json's FULL sizes through size-independent properties (the oracle finishes only small cases).

  cfg 2  2^20 envs single step      -- a strided sample of envs is compared with the C oracle teacher-forced, and
                                        step counters / done flags obey the TimeLimit + auto-reset identities
  cfg 3  2^26 envs, sharded          -- any split of the id range reproduces the unsplit run bit for bit (2^22 here x 16
                                        shards of the same code path; the full 2^26 run is bench.py's step_f32_64M)
  cfg 4  2^24 envs x 500 fused steps -- step conservation: every env executes exactly K steps, so
                                        sum(finished episode lengths) + sum(elapsed) == N*K + sum(elapsed before);
                                        episodes == new episode count; rollout(K1) + rollout(K2) == rollout(K1+K2)
  cfg 5  2^24 x 30 sampler           -- per-dim moments against the closed forms, support, determinism
"""
import math

import numpy as np
import pytest
import torch

import random_envs_b200 as random_envs
from oracle import c_oracle, dr_port

pytestmark = pytest.mark.gpu

SEARCH = [2.0, 20.0, 0.5, 3.0, 0.05, 0.3, 0.1, 1.0]


def _env(n, dtype, **kw):
    env = random_envs.RandomCartPoleVecEnv(n, dtype=dtype, **kw)
    env.set_dr_distribution("uniform", SEARCH); env.set_dr_training(True)
    return env


def test_cfg2_one_million_envs_step_sampled_against_the_oracle_and_counter_identities():
    n = 1 << 20
    env = _env(n, "float64", seed=3, max_episode_steps=30)
    env.reset()
    pick_np = np.arange(0, n, 4099)
    pick = torch.as_tensor(pick_np, device="cuda")
    total_done = 0
    for k in range(40):
        before = env.state.clone()[pick].cpu().numpy()
        xi = env.get_task()[pick].cpu().numpy()
        el_before = env.elapsed.clone()
        a = env.sample_actions().clone()
        obs, rew, done, info = env.step(a)
        ref = np.ascontiguousarray(before.T).copy()
        term = c_oracle.step_batch(ref, np.ascontiguousarray(xi.T), a[pick].cpu().numpy(), True)
        d = done[pick].cpu().numpy()
        trunc = info["TimeLimit.truncated"][pick].cpu().numpy()
        assert np.array_equal(d, term | (el_before[pick].cpu().numpy() + 1 >= 30))
        assert np.array_equal(trunc, ~term & d)
        alive = ~d
        assert np.max(np.abs(obs[pick].cpu().numpy()[alive] - ref.T[alive])) <= 1e-12
        assert bool((rew == 1).all())
        assert bool((env.elapsed[done] == 0).all()) and bool((env.elapsed[~done] == el_before[~done] + 1).all())
        total_done += int(done.sum())
    assert int(env.episode.sum()) == n + total_done


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_cfg3_sixteen_shards_equal_the_unsplit_run(dtype):
    n, shards = 1 << 22, 16
    mk = lambda m, id0: _truncnorm(random_envs.RandomCartPoleVecEnv(m, dtype=dtype, seed=8, env_id0=id0))
    whole = mk(n, 0)
    parts = [mk(n // shards, r * (n // shards)) for r in range(shards)]
    whole.reset(); [p.reset() for p in parts]
    w = (0.1, 0.1, 1.0, 0.3)
    for _ in range(3):
        a = whole.sample_actions().clone()
        whole.step(a)
        for r, p in enumerate(parts):
            p.step(a[r * (n // shards):(r + 1) * (n // shards)].clone())
    whole.rollout(w, 0.0, 40); [p.rollout(w, 0.0, 40) for p in parts]
    assert torch.equal(whole.state, torch.cat([p.state for p in parts]))
    assert torch.equal(whole.get_task(), torch.cat([p.get_task() for p in parts]))
    assert torch.equal(whole.elapsed, torch.cat([p.elapsed for p in parts]))
    combined = random_envs.combine_stats(np.stack([p.stats_tensor.cpu().numpy() for p in parts]))
    assert np.array_equal(combined, whole.stats_tensor.cpu().numpy())


def _truncnorm(env):
    env.set_dr_distribution("truncnorm", [9.8, 0.98, 1.0, 0.1, 0.2, 0.02, 0.5, 0.05]); env.set_dr_training(True)
    return env


@pytest.mark.parametrize("dtype,w", [("float32", (0.1, 0.1, 1.0, 0.3)), ("float32", (0.0, 0.0, 1.0, 0.0)),
                                     ("float64", (0.0, 0.0, 1.0, 0.0))])
def test_cfg4_sixteen_million_envs_500_fused_steps_conserve_steps(dtype, w):
    n, K = 1 << 24, 500
    env = _env(n, dtype, seed=2)
    env.reset()
    env.rollout(w, 0.0, 7)                                    # a mix of episode ages before the measured launch
    env.reset_stats()
    el0 = int(env.elapsed.sum(dtype=torch.int64))
    ep0 = int(env.episode.sum(dtype=torch.int64))
    env.rollout(w, 0.0, K)
    st = env.stats_tensor.cpu().numpy()
    el1 = int(env.elapsed.sum(dtype=torch.int64))
    ep1 = int(env.episode.sum(dtype=torch.int64))
    assert int(st[5]) + el1 == n * K + el0                    # every env executed exactly K steps
    assert int(st[0]) == ep1 - ep0 and st[1] == st[5]         # one statistic per finished episode; return == length
    assert 1 <= st[3] <= st[4] <= 500 and st[2] >= st[1] ** 2 / max(st[0], 1) * (1 - 1e-12)
    assert bool((env.elapsed < 500).all()) and bool((env.elapsed >= 0).all())
    assert bool((env.state[:, 0].abs() <= 2.4).all()) and bool((env.state[:, 2].abs() <= env.theta_threshold_radians).all())
    if w[0] > 0:
        assert st[1] / st[0] > 450                            # the stabilising policy: ~99.6 % of episodes reach 500


def test_cfg4_rollout_is_additive_in_K():
    n = 1 << 20
    a, b = _env(n, "float32", seed=6), _env(n, "float32", seed=6)
    a.reset(); b.reset()
    w = (0.0, 0.0, 1.0, 0.0)
    a.rollout(w, 0.0, 123); a.rollout(w, 0.0, 377)
    b.rollout(w, 0.0, 500)
    assert torch.equal(a.state, b.state) and torch.equal(a.get_task(), b.get_task()) and torch.equal(a.elapsed, b.elapsed)
    assert torch.equal(a.episode, b.episode)
    assert np.array_equal(a.stats_tensor.cpu().numpy(), b.stats_tensor.cpu().numpy())


@pytest.mark.parametrize("dr_type", ["uniform", "gaussian", "truncnorm"])
def test_cfg5_sixteen_million_humanoid_samples_moments_and_support(dr_type):
    n = 1 << 24
    nu = np.array(random_envs.HUMANOID_NOMINAL)
    s = random_envs.TaskSampler("RandomHumanoid-v0"); s.seed_dr(12)
    distr = np.stack([0.5 * nu, 1.5 * nu], 1).reshape(-1) if dr_type == "uniform" else np.stack([nu, 0.1 * nu], 1).reshape(-1)
    s.set_dr_distribution(dr_type, list(distr))
    x = s.sample_tasks_tensor(n, dtype=torch.float32)
    mean = x.double().mean(0).cpu().numpy()
    var = x.double().var(0).cpu().numpy()
    if dr_type == "uniform":
        want_mean, want_var = nu, (nu ** 2) / 12.0
        assert bool((x >= torch.tensor(0.5 * nu, device="cuda", dtype=torch.float32)).all())
        assert bool((x <= torch.tensor(1.5 * nu, device="cuda", dtype=torch.float32)).all())
    elif dr_type == "gaussian":
        want_mean, want_var = nu, (0.1 * nu) ** 2
        assert bool((x >= 0.1).all())
    else:
        want_mean, want_var = nu, dr_port.TN_VAR * (0.1 * nu) ** 2
        lo = torch.tensor(0.8 * nu, device="cuda", dtype=torch.float32); hi = torch.tensor(1.2 * nu, device="cuda", dtype=torch.float32)
        assert bool((x >= lo * (1 - 1e-6)).all()) and bool((x <= hi * (1 + 1e-6)).all())
    se = np.sqrt(want_var / n)
    assert np.all(np.abs(mean - want_mean) < 6 * se + 2e-7 * nu), np.max(np.abs(mean - want_mean) / se)
    assert np.all(np.abs(var / want_var - 1) < 6 * math.sqrt(2.0 / n) + 1e-4)
    s2 = random_envs.TaskSampler("RandomHumanoid-v0"); s2.seed_dr(12); s2.set_dr_distribution(dr_type, list(distr))
    assert torch.equal(s2.sample_tasks_tensor(n, dtype=torch.float32), x)
    s.check_dr_violations()


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("integrator,key", [("euler", "euler"), ("semi-implicit-euler", "semi_implicit")])
def test_episode_length_law_matches_the_live_references_demo_loop(golden_dir, dtype, integrator, key):
    """End to end: dynamics + termination + TimeLimit + reset + uniform DR resample under a random policy.  The golden
    file holds 20 000 episode lengths produced by the UNMODIFIED reference classes running their own demo loop
    (oracle/make_golden.py: set_random_task / reset / step(action_space.sample())); the RNG streams differ, so the
    comparison is distributional: two-sample KS at alpha = 0.001 plus mean and variance within 5 standard errors."""
    import os
    ref = np.load(os.path.join(golden_dir, "cartpole_episode_lengths.npz"))[key].astype(np.float64)
    n = 1 << 17
    env = random_envs.RandomCartPoleVecEnv(n, dtype=dtype, seed=77, kinematics_integrator=integrator)
    env.set_dr_distribution("uniform", SEARCH); env.set_dr_training(True)
    env.reset()
    first = torch.zeros(n, dtype=torch.int32, device="cuda")
    for k in range(500):
        el = env.elapsed.clone()
        _, _, done, _ = env.step(env.sample_actions())
        newly = done & (first == 0)
        first[newly] = el[newly] + 1
        if bool((first > 0).all()):
            break
    got = first.cpu().numpy().astype(np.float64)
    assert got.min() >= 1
    # two-sample KS
    grid = np.arange(1, 501)
    cdf_g = np.searchsorted(np.sort(got), grid, side="right") / got.size
    cdf_r = np.searchsorted(np.sort(ref), grid, side="right") / ref.size
    d = np.max(np.abs(cdf_g - cdf_r))
    crit = 1.95 * math.sqrt((got.size + ref.size) / (got.size * ref.size))
    assert d < crit, (d, crit)
    se_mean = math.sqrt(ref.var() / ref.size + got.var() / got.size)
    assert abs(got.mean() - ref.mean()) < 5 * se_mean, (got.mean(), ref.mean())
    assert abs(got.std() / ref.std() - 1) < 0.05
