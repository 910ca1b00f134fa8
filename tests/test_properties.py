"""T2 property tests (hypothesis).  CPU part: host logic; GPU part: kernel invariants over random shapes/seeds."""
import numpy as np
import pytest
from hypothesis import HealthCheck, assume, given, settings, strategies as st

import random_envs_b200 as random_envs
from oracle import c_oracle, cartpole_port as port

finite = st.floats(min_value=-1e3, max_value=1e3, allow_nan=False, allow_infinity=False)


@settings(max_examples=200, deadline=None)
@given(st.integers(1, 1 << 40), st.integers(1, 64))
def test_shard_ranges_tile_the_id_space(total, world):
    spans = [random_envs.shard_range(total, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == total
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1 and sorted(sizes, reverse=True) == sizes


@settings(max_examples=100, deadline=None)
@given(st.lists(finite, min_size=0, max_size=8), st.sampled_from(["uniform", "truncnorm", "gaussian"]))
def test_interleaved_distribution_prefix_rule(distr, dr_type):
    """random_env.py:102-121: pairs (a_i, b_i) fill a prefix; an odd trailing value is ignored."""
    env = random_envs.RandomCartPoleEnv()
    env.set_dr_distribution(dr_type, distr)
    k = len(distr) // 2
    first, second = (env.min_task, env.max_task) if dr_type == "uniform" else (env.mean_task, env.stdev_task)
    assert list(first[:k]) == distr[0:2 * k:2] and list(second[:k]) == distr[1:2 * k:2]
    assert not first[k:].any() and not second[k:].any() and env.sampling == dr_type
    cfg = env.dr_config()
    assert cfg.dim == 4 and [cfg.a[i] for i in range(4)] == list(first)


@settings(max_examples=300, deadline=None)
@given(st.tuples(st.floats(-3, 3), st.floats(-4, 4), st.floats(-0.3, 0.3), st.floats(-4, 4)),
       st.tuples(st.floats(2, 20), st.floats(0.5, 3), st.floats(0.05, 0.3), st.floats(0.1, 1.0)),
       st.integers(0, 1), st.booleans())
def test_port_and_c_oracle_agree_everywhere(state, xi, action, euler):
    """The two restatements of random_cartpole.py:176-205 are bit-identical on arbitrary inputs."""
    got, term = port.dynamics_step(state, xi, action, euler)
    st_c = np.array(state, np.float64).reshape(4, 1).copy()
    term_c = c_oracle.step_batch(st_c, np.array(xi).reshape(4, 1), np.array([action], np.uint8), euler)
    assert tuple(st_c[:, 0]) == got and bool(term_c[0]) == term
    x, th = got[0], got[2]
    assert term == (abs(x) > 2.4 or abs(th) > port.THETA_THRESHOLD)       # strict inequalities on the NEW state


# ---------------------------------------------------------------------------------------------- GPU invariants
gpu = pytest.mark.gpu


@gpu
@settings(max_examples=12, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(st.integers(1, 5000), st.integers(0, 2 ** 63 - 1), st.integers(1, 40), st.sampled_from(["float32", "float64"]),
       st.booleans())
def test_step_invariants_any_size_seed_limit(n, seed, limit, dtype, euler):
    """reward == 1; done <=> (outside thresholds or elapsed hit the limit); counters; obs of done envs is fresh."""
    import torch
    env = random_envs.RandomCartPoleVecEnv(n, dtype=dtype, seed=seed, max_episode_steps=limit,
                                           kinematics_integrator="euler" if euler else "semi")
    env.set_dr_distribution("uniform", [2.0, 20.0, 0.5, 3.0, 0.05, 0.3, 0.1, 1.0]); env.set_dr_training(True)
    env.reset()
    for _ in range(min(2 * limit, 50)):
        el_before = env.elapsed.clone()
        obs, rew, done, info = env.step(env.sample_actions())
        trunc = info["TimeLimit.truncated"]
        assert bool((rew == 1).all())
        assert bool((~trunc | done).all())                                  # truncated => done
        hit = el_before + 1 >= limit
        assert bool((trunc <= hit).all()) and bool((hit <= done).all())     # truncation only at the limit; limit => done
        assert bool((env.elapsed[done] == 0).all()) and bool((env.elapsed[~done] == el_before[~done] + 1).all())
        assert bool((obs[done].abs() <= 0.05).all())
        inside = (obs[:, 0].abs() <= 2.4) & (obs[:, 2].abs() <= env.theta_threshold_radians)
        assert bool(inside[~done].all())
        assert bool((env.elapsed < limit).all())


@gpu
@settings(max_examples=8, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(st.integers(2, 3000), st.integers(0, 2 ** 32), st.integers(1, 2999), st.sampled_from(["float32", "float64"]))
def test_any_split_of_the_env_range_gives_identical_trajectories(n, seed, cut, dtype):
    import torch
    cut = min(cut, n - 1)
    mk = lambda m, id0: _mk(m, id0, seed, dtype)
    whole, left, right = mk(n, 0), mk(cut, 0), mk(n - cut, cut)
    for e in (whole, left, right):
        e.reset()
    for _ in range(25):
        ow, _, dw, _ = whole.step(whole.sample_actions())
        ol, _, dl, _ = left.step(left.sample_actions())
        o_r, _, dr, _ = right.step(right.sample_actions())
        assert torch.equal(ow, torch.cat([ol, o_r])) and torch.equal(dw, torch.cat([dl, dr]))
    assert torch.equal(whole.get_task(), torch.cat([left.get_task(), right.get_task()]))


def _mk(n, env_id0, seed, dtype):
    env = random_envs.RandomCartPoleVecEnv(n, dtype=dtype, seed=seed, env_id0=env_id0, max_episode_steps=17)
    env.set_dr_distribution("truncnorm", [9.8, 1.0, 1.0, 0.1, 0.1, 0.02, 0.5, 0.05]); env.set_dr_training(True)
    return env


@gpu
@settings(max_examples=10, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(st.sampled_from(sorted(random_envs.XI_TABLES)), st.sampled_from(["uniform", "truncnorm", "gaussian"]),
       st.integers(1, 3000), st.integers(0, 2 ** 40))
def test_sampler_outputs_respect_their_support(env_id, dr_type, n, seed):
    table = random_envs.XI_TABLES[env_id]
    lo = np.array([b[0] for b in table.search_bounds]); hi = np.array([b[1] for b in table.search_bounds])
    lb = np.array(table.lower_bounds)
    s = random_envs.TaskSampler(env_id); s.seed_dr(seed)
    mid, sd = (lo + hi) / 2, (hi - lo) / 10
    distr = np.stack([lo, hi], 1).reshape(-1) if dr_type == "uniform" else np.stack([mid, sd], 1).reshape(-1)
    s.set_dr_distribution(dr_type, list(distr))
    x = s.sample_tasks(n)
    assert x.shape == (n, len(lo)) and np.all(np.isfinite(x))
    if dr_type == "uniform":
        assert np.all(x >= lo) and np.all(x <= hi)
    elif dr_type == "truncnorm":
        assert np.all(x >= np.maximum(lb, mid - 2 * sd) - 1e-12) and np.all(x <= mid + 2 * sd + 1e-12)
    else:
        assert np.all(x >= 0.1)
    assert np.array_equal(x, _again(env_id, dr_type, distr, n, seed))       # deterministic in (seed, call)


def _again(env_id, dr_type, distr, n, seed):
    s = random_envs.TaskSampler(env_id); s.seed_dr(seed)
    s.set_dr_distribution(dr_type, list(distr))
    return s.sample_tasks(n)


@gpu
@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.filter_too_much])
@given(st.integers(1, 700), st.integers(1, 70), st.integers(0, 12), st.sampled_from(["float32", "float64"]),
       st.sampled_from([(0.0, 0.0, 1.0, 0.0), (0.0, 0.0, -1.0, 0.0), (0.1, 0.1, 1.0, 0.3), (0.0, 0.0, 0.0, 0.0)]),
       st.sampled_from([-1.0, 0.0, 1.0]), st.integers(0, 2 ** 40), st.booleans())
def test_fused_rollout_equals_single_steps_for_any_shape(n, K, limit, dtype, w, b, seed, euler):
    """The deferred, warp-batched resets and the packed env-pair kernel must not change a single bit: any n (odd sizes
    leave a half-empty pair), any K (incl. K = 1 and episodes ending on the last step), any step limit (limit 1 ends
    every episode at every step), destabilising / stabilising / constant policies."""
    import torch
    nz = [c for c in range(4) if w[c] != 0.0]
    # the fp32 kernel evaluates w.s + b as an FMA chain; torch reproduces that bit for bit only for tie-free forms
    assume(dtype == "float64" or not nz or (len(nz) == 1 and b == 0.0))
    mk = lambda: _mk_roll(n, seed, dtype, limit, euler)
    fused, stepped = mk(), mk()
    fused.reset(); stepped.reset()
    fused.rollout(w, b, K)
    ends = 0
    for _ in range(K):
        s = stepped.state
        if not nz:
            act = torch.full((n,), int(b > 0), dtype=torch.uint8, device="cuda")
        elif dtype == "float32":
            act = (s[:, nz[0]] * w[nz[0]] > 0).to(torch.uint8)
        else:                       # policy_action<double>: left-to-right, separately rounded
            acc = s[:, 0] * w[0]
            for c in range(1, 4):
                acc = acc + s[:, c] * w[c]
            act = (acc + b > 0).to(torch.uint8)
        _, _, done, _ = stepped.step(act)
        ends += int(done.sum())
    assert torch.equal(fused.state, stepped.state) and torch.equal(fused.get_task(), stepped.get_task())
    assert torch.equal(fused.elapsed, stepped.elapsed) and torch.equal(fused.episode, stepped.episode)
    assert int(fused.stats_tensor[0]) == ends


def _mk_roll(n, seed, dtype, limit, euler):
    env = random_envs.RandomCartPoleVecEnv(n, dtype=dtype, seed=seed, max_episode_steps=limit,
                                           kinematics_integrator="euler" if euler else "semi")
    env.set_dr_distribution("uniform", [2.0, 20.0, 0.5, 3.0, 0.05, 0.3, 0.1, 1.0]); env.set_dr_training(True)
    return env
