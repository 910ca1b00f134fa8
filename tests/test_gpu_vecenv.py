"""T2/T4/T5 on the GPU: vector-env semantics, fused rollout, sharding invariance, the gym drop-in."""
import math

import numpy as np
import pytest
import torch

import random_envs_b200 as random_envs
from random_envs_b200 import gym
from oracle import c_oracle, cartpole_port as port

pytestmark = pytest.mark.gpu

SEARCH = [2.0, 20.0, 0.5, 3.0, 0.05, 0.3, 0.1, 1.0]
LO = np.array(SEARCH[0::2]); HI = np.array(SEARCH[1::2])


def _np(t):
    return t.detach().cpu().numpy()


def _dr_env(n, dtype="float32", seed=0, **kw):
    env = random_envs.RandomCartPoleVecEnv(n, dtype=dtype, seed=seed, **kw)
    env.set_dr_distribution("uniform", SEARCH)
    env.set_dr_training(True)
    return env


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_reset_and_step_contract(dtype):
    n = 10007
    env = _dr_env(n, dtype)
    obs = env.reset()
    tdt = torch.float32 if dtype == "float32" else torch.float64
    assert obs.shape == (n, 4) and obs.dtype == tdt and obs.is_cuda
    o = _np(obs)
    assert np.all(o >= -0.05) and np.all(o <= 0.05) and np.abs(o.mean()) < 1e-3
    xi = _np(env.get_task())
    assert np.all(xi >= LO) and np.all(xi <= HI)
    assert np.array_equal(_np(env.episode), np.ones(n)) and np.array_equal(_np(env.elapsed), np.zeros(n))
    obs, rew, done, info = env.step(env.sample_actions())
    assert rew.shape == (n,) and rew.dtype == tdt and done.dtype == torch.bool and done.shape == (n,)
    assert info["TimeLimit.truncated"].dtype == torch.bool and not bool(info["TimeLimit.truncated"].any())
    assert bool((rew == 1).all())
    # accepted action containers: int64 tensor, numpy array, python list (n=3 env)
    small = _dr_env(3, dtype); small.reset()
    for a in (torch.tensor([0, 1, 1]), np.array([1, 0, 1]), [1, 1, 0], torch.tensor([True, False, True])):
        small.step(a)
    with pytest.raises(ValueError):
        small.step([0, 1])
    with pytest.raises(AssertionError, match="invalid"):
        small.step(torch.tensor([0.0, 1.0, 1.0]))
    # an action outside Discrete(2) is flagged by the step kernel and raised at the next synchronising call
    strict = random_envs.RandomCartPoleVecEnv(3); strict.reset()
    strict.step([0, 2, 1])
    with pytest.raises(AssertionError, match="invalid"):
        strict.check_dr_violations()


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_auto_reset_properties_under_random_policy(dtype):
    """T2 properties on 2^18 envs x 60 steps: reward, strict thresholds, counters, xi resample only on reset."""
    n = 1 << 18
    env = _dr_env(n, dtype, seed=2)
    env.reset()
    total_done = 0
    for k in range(60):
        xi_before = env.get_task().clone()
        ep_before = env.episode.clone()
        el_before = env.elapsed.clone()
        obs, rew, done, info = env.step(env.sample_actions())
        assert bool((rew == 1).all())
        o = obs
        inside = (o[:, 0].abs() <= 2.4) & (o[:, 2].abs() <= env.theta_threshold_radians)
        assert bool(inside[~done].all())                                   # not done => inside thresholds
        fresh = (o.abs() <= 0.05).all(1)
        assert bool(fresh[done].all())                                     # done => obs is the reset obs
        assert bool((env.elapsed[done] == 0).all()) and bool((env.elapsed[~done] == el_before[~done] + 1).all())
        assert bool((env.episode[done] == ep_before[done] + 1).all()) and bool((env.episode[~done] == ep_before[~done]).all())
        xi_after = env.get_task()
        assert bool((xi_after[~done] == xi_before[~done]).all())           # xi untouched while the episode runs
        changed = (xi_after[done] != xi_before[done]).any(1)
        assert bool(changed.all())                                         # resampled on reset
        total_done += int(done.sum())
    xi = _np(env.get_task()).astype(np.float64)
    assert np.all(xi >= LO - 1e-6) and np.all(xi <= HI + 1e-6)
    mean_len = n * 60 / max(total_done, 1)
    assert 15 < mean_len < 45, mean_len                                    # SURVEY: ~27 under a random policy
    # without dr_training xi stays put across resets
    env.set_dr_training(False)
    xi0 = env.get_task().clone()
    for k in range(40):
        env.step(env.sample_actions())
    assert bool((env.get_task() == xi0).all())


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_sharding_invariance(dtype):
    """Trajectories depend on the GLOBAL env id only: one 6000-env shard == shards of 1000 + 5000."""
    n, K = 6000, 80
    whole = _dr_env(n, dtype, seed=9)
    a_part = _dr_env(1000, dtype, seed=9, env_id0=0)
    b_part = _dr_env(5000, dtype, seed=9, env_id0=1000)
    for e in (whole, a_part, b_part):
        e.reset()
    for k in range(K):
        ow, rw, dw, _ = whole.step(whole.sample_actions())
        oa, ra, da, _ = a_part.step(a_part.sample_actions())
        ob, rb, db, _ = b_part.step(b_part.sample_actions())
        assert torch.equal(ow, torch.cat([oa, ob])) and torch.equal(dw, torch.cat([da, db])), k
    assert torch.equal(whole.get_task(), torch.cat([a_part.get_task(), b_part.get_task()]))
    assert torch.equal(whole.episode, torch.cat([a_part.episode, b_part.episode]))
    lo, hi = random_envs.shard_range(n, 1, 6)
    assert (lo, hi) == (1000, 2000)


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("w,b", [((0.0, 0.0, 1.0, 0.0), 0.0), ((0.0, 0.0, 0.0, 0.0), 1.0), ((0.0, 0.0, 0.0, 0.0), -1.0)])
def test_rollout_equals_repeated_step(dtype, w, b):
    """The fused K-step kernel and K single-step launches share one device function: bit-identical."""
    n, K = 5000, 150
    fused = _dr_env(n, dtype, seed=4); stepped = _dr_env(n, dtype, seed=4)
    fused.reset(); stepped.reset()
    fused.rollout(w, b, K)
    episodes = 0
    ret_sum = 0.0
    for k in range(K):
        s = stepped.obs
        act = ((s[:, 2] * w[2] + b) > 0).to(torch.uint8) if w[2] != 0 else torch.full((n,), int(b > 0), dtype=torch.uint8, device="cuda")
        el = stepped.elapsed.clone()
        _, _, done, _ = stepped.step(act)
        episodes += int(done.sum()); ret_sum += float((el[done] + 1).sum())
    assert torch.equal(fused.obs, stepped.obs) and torch.equal(fused.get_task(), stepped.get_task())
    assert torch.equal(fused.elapsed, stepped.elapsed) and torch.equal(fused.episode, stepped.episode)
    st = fused.episode_stats()
    assert st["episodes"] == episodes and abs(st["mean_return"] * episodes - ret_sum) < 1e-6
    # two half-length rollouts == one full-length rollout
    halves = _dr_env(n, dtype, seed=4); halves.reset()
    halves.rollout(w, b, K // 2); halves.rollout(w, b, K - K // 2)
    assert torch.equal(halves.obs, fused.obs) and torch.equal(halves.stats_tensor, fused.stats_tensor)


def test_fp64_rollout_vs_c_oracle_closed_loop():
    """Closed loop with the in-kernel linear policy vs the oracle's left-to-right dot product.

    States are compared after 30 steps, while the unstable mode's amplification (e^{lambda t}, lambda up to ~13/s
    for short poles in strong gravity) keeps a 1-ulp sin/cos difference below 1e-9; counters must agree exactly
    there.  The remaining 470 steps are compared
    through the episode statistics (trajectories legitimately decorrelate beyond ~350 steps)."""
    n, seed = 30000, 13
    w, b = (0.1, 0.1, 1.0, 0.3), 0.0
    env = _dr_env(n, "float64", seed=seed)
    obs0 = _np(env.reset()).copy(); xi0 = _np(env.get_task()).copy()
    st = np.ascontiguousarray(obs0.T); xi = np.ascontiguousarray(xi0.T)
    el = np.zeros(n, np.int32); ep = np.ones(n, np.uint32)
    out = c_oracle.closed_loop(st, xi, el, ep, seed, 0, 1, 30, max_steps=500, w=np.array(w), b=b, lo=LO, hi=HI)
    env.rollout(w, b, 30)
    same = (_np(env.episode).astype(np.uint32) == ep) & (_np(env.elapsed) == el)
    assert same.mean() >= 0.9999, same.mean()      # a near-tie |w.s| ~ 1e-16 may flip one action in ~1e5 envs
    assert np.max(np.abs(_np(env.obs)[same] - st.T[same])) <= 1e-9
    assert np.array_equal(_np(env.get_task())[same], xi.T[same])
    got = _np(env.stats_tensor).copy()
    assert abs(got[0] - out["stats"][0]) <= 3 and abs(got[1] - out["stats"][1]) <= 300
    out = c_oracle.closed_loop(st, xi, el, ep, seed, 0, 31, 470, max_steps=500, w=np.array(w), b=b, lo=LO, hi=HI,
                               stats=out["stats"])
    env.rollout(w, b, 470)
    got = _np(env.stats_tensor); want = out["stats"]
    assert abs(got[0] / want[0] - 1) < 0.01 and abs(got[1] / got[0] - want[1] / want[0]) < 5.0
    assert got[4] == 500.0 and want[4] == 500.0 and got[3] >= 1.0
    assert (_np(env.episode) == 2).mean() > 0.95   # SURVEY: 99.6 % of DR envs survive 500 steps under this policy


def test_fp32_rollout_statistics_match_fp64():
    n, K = 1 << 17, 500
    out = {}
    for dtype in ("float32", "float64"):
        env = _dr_env(n, dtype, seed=6); env.reset()
        env.rollout((0.0, 0.0, 1.0, 0.0), 0.0, K)          # theta bang-bang: mean episode length ~78 (SURVEY cfg 4)
        out[dtype] = env.episode_stats()
    a, b = out["float32"], out["float64"]
    assert a["episodes"] > n and abs(a["episodes"] / b["episodes"] - 1) < 0.01
    assert abs(a["mean_return"] / b["mean_return"] - 1) < 0.01 and 40 < b["mean_return"] < 150
    assert a["max_return"] <= 500 and a["min_return"] >= 1


def test_truncnorm_dr_in_step_and_rollout():
    """BASELINE config 3's distribution: truncnorm around nominal, pole_mass mean at the lower bound."""
    n = 1 << 16
    env = random_envs.RandomCartPoleVecEnv(n, dtype="float32", seed=3)
    env.set_dr_distribution("truncnorm", [9.8, 0.98, 1.0, 0.1, 0.1, 0.02, 0.5, 0.05])
    env.set_dr_training(True)
    env.reset()
    xi = _np(env.get_task()).astype(np.float64)
    assert abs(np.mean(xi[:, 2] == np.float32(0.1)) - 0.125) < 0.01        # p^3 point mass at lb
    assert np.all(np.abs(xi[:, 0] - 9.8) <= 2 * 0.98 + 1e-5) and np.all(xi[:, 2] >= 0.0999999)
    env.rollout((0.1, 0.1, 1.0, 0.3), 0.0, 300)
    env.check_dr_violations()
    xi2 = _np(env.get_task()).astype(np.float64)
    assert np.all(np.abs(xi2[:, 3] - 0.5) <= 0.1 + 1e-6)
    # gaussian DR that cannot satisfy the 0.1 floor is reported, not silently accepted
    bad = random_envs.RandomCartPoleVecEnv(64, seed=1)
    bad.set_dr_distribution("gaussian", [9.8, 1.0, 1.0, 0.1, -3.0, 0.1, 0.5, 0.05]); bad.set_dr_training(True)
    bad.reset()
    with pytest.raises(Exception, match="Not all samples were above > 0.1 after 2 attempts"):
        bad.check_dr_violations()


def test_masked_reset_set_task_and_checkpoint_roundtrip():
    n = 1024
    env = _dr_env(n, "float64", seed=8); env.reset()
    for _ in range(5):
        env.step(env.sample_actions())
    before = env.obs.clone(); xi_before = env.get_task().clone()
    mask = torch.zeros(n, dtype=torch.bool, device="cuda"); mask[::3] = True
    env.reset(mask)
    assert torch.equal(env.obs[~mask], before[~mask]) and bool((env.obs[mask] != before[mask]).any(1).all())
    assert bool((env.get_task()[mask] != xi_before[mask]).any(1).all()) and torch.equal(env.get_task()[~mask], xi_before[~mask])
    env.set_task(5.0, 2.0, 0.2, 0.8)
    assert bool((env.get_task() == torch.tensor([5.0, 2.0, 0.2, 0.8], device="cuda", dtype=torch.float64)).all())
    env.set_random_task()
    assert np.all(_np(env.get_task()) >= LO) and len(np.unique(_np(env.get_task())[:, 0])) > n // 2
    sd = env.state_dict()
    twin = _dr_env(n, "float64", seed=999); twin.load_state_dict(sd)
    for _ in range(30):
        a = env.sample_actions().clone()
        o1, _, d1, _ = env.step(a); o2, _, d2, _ = twin.step(a)
        assert torch.equal(o1, o2) and torch.equal(d1, d2)


def test_large_env_ids_and_ticks_follow_the_oracle_contract():
    """Global env ids beyond 2^32 (shards of a huge job) and clocks beyond 2^32 steps key distinct streams."""
    base = (1 << 33) + 12345
    for dtype, npdt in (("float64", np.float64), ("float32", np.float32)):
        env = _dr_env(64, dtype, seed=17, env_id0=base)
        env._tick = (1 << 40) + 7
        obs = _np(env.reset()); xi = _np(env.get_task())
        for i in (0, 63):
            assert np.array_equal(obs[i], c_oracle.init_state(17, base + i, (1 << 40) + 7, npdt))
            assert np.array_equal(xi[i], c_oracle.xi_uniform(17, base + i, (1 << 40) + 7, LO, HI, dtype=npdt))
        a = env.sample_actions()
        assert np.array_equal(_np(a), c_oracle.random_actions(64, base, 17, ((1 << 40) + 8) & 0xFFFFFFFF))


def test_step_host_matches_step():
    n = 4099
    dev = _dr_env(n, "float32", seed=5); host = _dr_env(n, "float32", seed=5)
    dev.reset(); host.reset()
    for k in range(10):
        a = dev.sample_actions().clone()
        o, r, d, info = dev.step(a)
        ho, hr, hd, ht = host.step_host(_np(a))
        assert np.array_equal(ho, _np(o)) and np.array_equal(hr, _np(r)) and np.array_equal(hd, _np(d))
        assert np.array_equal(ht, _np(info["TimeLimit.truncated"]))


def test_step_host_without_auto_reset_carries_the_zero_rewards():
    """Only with auto-reset is the reward a host-side constant; the steps-beyond-done rule must cross PCIe."""
    n = 515
    dev = random_envs.RandomCartPoleVecEnv(n, dtype="float64", seed=9, auto_reset=False, max_episode_steps=0)
    host = random_envs.RandomCartPoleVecEnv(n, dtype="float64", seed=9, auto_reset=False, max_episode_steps=0)
    dev.reset(); host.reset()
    zeros = 0
    for k in range(60):
        a = torch.ones(n, dtype=torch.uint8, device="cuda")            # constant push: everyone falls, then keeps stepping
        o, r, d, _ = dev.step(a)
        ho, hr, hd, _ = host.step_host(_np(a))
        assert np.array_equal(hr, _np(r)) and np.array_equal(hd, _np(d)) and np.array_equal(ho, _np(o))
        zeros += int((hr == 0).sum())
    assert zeros > 0
    assert host.host_bytes_per_step()[1] > dev.num_envs * (4 * 8 + 1 + 8)


# ------------------------------------------------------------------------------------------------ drop-in
def test_dropin_random_policy_loop_like_the_reference_demo():
    """test_random_policy.py:12-32 without rendering, plus README.md:52-66."""
    env = gym.make("RandomCartPole-v0")
    env.seed(0)
    env.set_dr_distribution(dr_type="uniform", distr=SEARCH)
    env.set_dr_training(True)
    state = env.reset()
    assert isinstance(state, np.ndarray) and state.shape == (4,) and state.dtype == np.float64
    assert np.all(np.abs(state) <= 0.05)
    tasks = [env.get_task().copy()]
    episodes, steps_in_ep, lengths = 0, 0, []
    for _ in range(600):
        state, reward, done, info = env.step(env.action_space.sample())
        steps_in_ep += 1
        assert isinstance(reward, float) and reward == 1.0 and isinstance(done, bool) and isinstance(info, dict)
        inside = abs(state[0]) <= 2.4 and abs(state[2]) <= env.theta_threshold_radians
        assert done == (not inside)
        if done:
            env.reset()
            tasks.append(env.get_task().copy()); lengths.append(steps_in_ep); episodes += 1; steps_in_ep = 0
    assert episodes >= 8 and 8 < np.mean(lengths) < 80
    t = np.array(tasks)
    assert len(np.unique(t[:, 0])) == len(t) and np.all(t >= LO) and np.all(t <= HI)   # resampled at every reset
    env.set_dr_training(False)
    fixed = env.get_task().copy(); env.reset()
    assert np.array_equal(env.get_task(), fixed)


def test_dropin_matches_oracle_port_step_for_step():
    env = random_envs.RandomCartPoleEnv()
    ref = port.CartPolePort()
    rs = np.random.RandomState(0)
    for episode in range(5):
        xi = LO + (HI - LO) * rs.uniform(size=4)
        s0 = rs.uniform(-0.05, 0.05, 4)
        env.reset(); env.set_task(*xi); env.state = tuple(s0); env.steps_beyond_done = None
        ref.set_task(*xi); ref.state = tuple(s0); ref.steps_beyond_done = None
        for k in range(120):
            a = int(rs.randint(2))
            o1, r1, d1, _ = env.step(a); o2, r2, d2, _ = ref.step(a)
            assert np.max(np.abs(o1 - o2)) <= 1e-12 * max(1.0, np.abs(o2).max()) and r1 == r2 and d1 == d2, (episode, k)
            assert env.steps_beyond_done == ref.steps_beyond_done
            if k > 60 and d1:
                break
    assert env.polemass_length == 0.05 and abs(env.total_mass - (xi[1] + xi[2])) < 1e-15


def test_dropin_timelimit_wrapper_truncates_at_500():
    env = gym.make("RandomCartPole-v0")
    env.seed(3); s = env.reset()
    w = np.array([0.1, 0.1, 1.0, 0.3])
    for k in range(500):
        s, r, done, info = env.step(int(w @ s > 0))
        if k < 499:
            assert not done
    assert done and info["TimeLimit.truncated"] is True


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("n,K,limit", [(1, 9, 500), (1000, 130, 500), (4099, 77, 11), (300, 600, 500)])
def test_random_policy_rollout_equals_step_with_sample_actions(dtype, n, K, limit):
    """rollout(w=None): the reference's demo loop (test_random_policy.py:25-32: action_space.sample(), reset on done)
    fused into one launch == K x step(sample_actions()), bit for bit (env ids beyond one 128-env Philox block too)."""
    mk = lambda: random_envs.RandomCartPoleVecEnv(n, dtype=dtype, seed=12, env_id0=(1 << 33) + 77, max_episode_steps=limit)
    fused, stepped = mk(), mk()
    for e in (fused, stepped):
        e.set_dr_distribution("uniform", SEARCH); e.set_dr_training(True); e.reset()
    fused.rollout(None, 0.0, K)
    ends = 0
    for _ in range(K):
        _, _, done, _ = stepped.step(stepped.sample_actions())
        ends += int(done.sum())
    assert torch.equal(fused.obs, stepped.obs) and torch.equal(fused.get_task(), stepped.get_task())
    assert torch.equal(fused.elapsed, stepped.elapsed) and torch.equal(fused.episode, stepped.episode)
    st = fused.episode_stats()
    assert st["episodes"] == ends
    if n >= 1000 and limit == 500:
        assert 15 < st["mean_return"] < 40                    # random policy over the search bounds: ~27 steps


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("pre,K", [(125, 1), (126, 1), (126, 2), (127, 1), (127, 130), (120, 300), (254, 4)])
def test_random_policy_rollout_across_action_block_boundaries(dtype, pre, K):
    """An env's 128 action bits run out at every multiple of 128 of the step clock; in the fused rollout the lane then
    parks for the warp's reset pass to fetch the next block.  Start the launch `pre` ticks into the stream so that the
    boundary falls on its first step, its last step, in the middle, twice, or right after the launch."""
    n = 777
    mk = lambda: random_envs.RandomCartPoleVecEnv(n, dtype=dtype, seed=5, env_id0=9000, max_episode_steps=40)
    fused, stepped = mk(), mk()
    for e in (fused, stepped):
        e.set_dr_distribution("uniform", SEARCH); e.set_dr_training(True); e.reset()
        for _ in range(pre):
            e.step(e.sample_actions())
    assert torch.equal(fused.obs, stepped.obs)
    fused.rollout(None, 0.0, K)
    for _ in range(K):
        stepped.step(stepped.sample_actions())
    assert torch.equal(fused.obs, stepped.obs) and torch.equal(fused.get_task(), stepped.get_task())
    assert torch.equal(fused.elapsed, stepped.elapsed) and torch.equal(fused.episode, stepped.episode)
    # and the stream continues identically after the launch (the clock advanced by K)
    a, b = fused.sample_actions().clone(), stepped.sample_actions().clone()
    assert torch.equal(a, b)


def test_in_place_edits_of_the_distribution_arrays_take_effect():
    """The reference reads min_task / max_task at every draw (random_env.py:151): editing them in place between steps
    changes what the next reset samples, without calling set_dr_distribution again."""
    env = random_envs.RandomCartPoleVecEnv(4096, dtype="float64", seed=3, max_episode_steps=1)   # every step resets
    env.set_dr_distribution("uniform", SEARCH); env.set_dr_training(True)
    env.reset()
    env.step(env.sample_actions())
    g = _np(env.get_task())[:, 0]
    assert g.min() >= 2.0 and g.max() <= 20.0 and g.max() > 15.0
    env.min_task[0], env.max_task[0] = 30.0, 31.0
    env.step(env.sample_actions())
    g = _np(env.get_task())[:, 0]
    assert g.min() >= 30.0 and g.max() <= 31.0


def test_resident_scalar_env_equals_the_launch_per_step_path_and_sample_task(monkeypatch):
    """The drop-in gym env is served by a resident kernel (renv_cartpole_scalar_serve); RENV_SCALAR_RESIDENT=0 is the
    launch-per-step path.  Same seed -> the same episodes bit for bit, including the xi drawn on reset, which equals
    what RandomEnv.sample_task() returns for the same call index."""
    def run(resident, dr_type, distr, noisy=False):
        monkeypatch.setenv("RENV_SCALAR_RESIDENT", "1" if resident else "0")
        env = random_envs.RandomCartPoleEnv(noisy=noisy)
        env.seed(5)
        env.set_dr_distribution(dr_type, distr); env.set_dr_training(True)
        rs = np.random.RandomState(1)
        out = [env.reset().copy(), env.get_task().copy()]
        for k in range(150):
            o, r, d, _ = env.step(int(rs.randint(2)))
            out.append(np.concatenate([o, [r, float(d)]]))
            if d and k % 3 == 0:
                out.append(env.reset().copy()); out.append(env.get_task().copy())
        env.close()
        return out
    cases = [("uniform", SEARCH, False), ("truncnorm", [9.8, 0.98, 1.0, 0.1, 0.11, 0.02, 0.5, 0.05], False),
             ("gaussian", [9.8, 0.98, 1.0, 0.1, 0.2, 0.02, 0.5, 0.05], True)]
    for dr_type, distr, noisy in cases:
        a, b = run(True, dr_type, distr, noisy), run(False, dr_type, distr, noisy)
        assert len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b)), dr_type
    monkeypatch.setenv("RENV_SCALAR_RESIDENT", "1")
    monkeypatch.setenv("RENV_SCALAR_LEASE_US", "300")
    e1, e2 = random_envs.RandomCartPoleEnv(), random_envs.RandomCartPoleEnv()
    for e in (e1, e2):
        e.seed(9); e.set_dr_distribution("uniform", SEARCH); e.set_dr_training(True)
    e1.reset()
    assert np.array_equal(e1.get_task(), e2.sample_task())           # call index 0 of the same Philox stream
    e1.reset()
    assert np.array_equal(e1.get_task(), e2.sample_task())           # call index 1
    # the lease: an idle env parks itself and is revived transparently
    import time
    s_before = e1.state
    time.sleep(0.01)
    o, r, d, _ = e1.step(1)
    assert e1._core.exited[0] != 0 and np.all(np.isfinite(o)) and e1.state != s_before
    with pytest.raises(Exception, match="Not all samples were above"):
        bad = random_envs.RandomCartPoleEnv()
        bad.set_dr_distribution("gaussian", [-5.0, 0.1, 1.0, 0.1, 0.1, 0.01, 0.5, 0.05]); bad.set_dr_training(True)
        bad.reset()


@pytest.mark.parametrize("noisy", [False, True])
def test_scalar_env_lookahead_equals_the_synchronous_protocol(monkeypatch, noisy):
    """The resident kernel publishes both outcomes of the next step (renv_scalar_ctrl.next) and env.step returns one of
    them without waiting for the round trip (RENV_SCALAR_LOOKAHEAD, default on).  The episodes equal those of the
    synchronous request/acknowledge protocol bit for bit -- with resets, state / task / steps_beyond_done edits between
    steps, stepping past `done`, a re-seed, and a lease that expires while a step is still in flight."""
    import time

    def run(lookahead):
        monkeypatch.setenv("RENV_SCALAR_LOOKAHEAD", "1" if lookahead else "0")
        monkeypatch.setenv("RENV_SCALAR_LEASE_US", "200")
        env = random_envs.RandomCartPoleEnv(noisy=noisy)
        assert env._core.lookahead is lookahead
        env.seed(3)
        env.set_dr_distribution("uniform", SEARCH); env.set_dr_training(True)
        rs = np.random.RandomState(2)
        out = [env.reset().copy(), env.get_task().copy()]
        for k in range(400):
            o, r, d, _ = env.step(int(rs.randint(2)))
            out.append(np.concatenate([o, [r, float(d)], env.state, [-1 if env.steps_beyond_done is None else env.steps_beyond_done]]))
            if d and k % 4 != 0:
                out.append(env.reset().copy()); out.append(env.get_task().copy())
            if k == 50:
                env.state = (0.01, -0.02, 0.03, 0.04)                  # user-assigned state between two steps
            if k == 90:
                env.set_task(9.0, 1.2, 0.15, 0.45)
            if k == 130:
                env.kinematics_integrator = "semi-implicit"
            if k == 170:
                time.sleep(0.005)                                      # the lease expires with the last step in flight
            if k == 200:
                env.seed(3); out.append(env.reset().copy())            # re-seeding restarts the stream
            if k == 260:
                env.noise_level = 4e-4
        env.close()
        return out
    a, b = run(True), run(False)
    assert len(a) == len(b)
    for i, (x, y) in enumerate(zip(a, b)):
        assert np.array_equal(x, y), i
