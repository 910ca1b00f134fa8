"""T1: CUDA step kernels vs the oracle, through the C ABI (ctypes -> librenv_b200.so).

Tolerances (stated here, as BASELINE.json north_star asks):
  fp64  |state - oracle| <= 1e-12 per step (teacher-forced: every step starts from the reference's state);
        free-running along whole episodes <= 1e-10 (golden, <= 110 steps) / 1e-8 (60-step DR episodes), because
        the pole's unstable mode multiplies a 1-ulp sin/cos difference by e^{lambda t}, lambda = 4..13 1/s;
        done flags / episode lengths / counters bit-exact.  (Expected per-step error: ~1e-16.)
  fp32  teacher-forced one-step error <= 5e-6 absolute at every one of 500 steps; free-running error
        <= 1e-5 @ step 1, 1e-4 @ 20 steps (the pole's unstable mode amplifies rounding by ~e^{4t}).
"""
import json
import os

import numpy as np
import pytest
import torch

import random_envs_b200 as random_envs
from oracle import c_oracle

pytestmark = pytest.mark.gpu

FP64_TOL = 1e-12
FP32_STEP_TOL = 5e-6


def _np(t):
    return t.detach().cpu().numpy()


def _vec(n, dtype="float64", **kw):
    kw.setdefault("auto_reset", False)
    kw.setdefault("max_episode_steps", 0)
    return random_envs.RandomCartPoleVecEnv(n, dtype=dtype, **kw)


def test_known_answers_fp64(golden_dir):
    rows = json.load(open(os.path.join(golden_dir, "cartpole_known_answers.json")))
    for r in rows:
        env = _vec(1, kinematics_integrator="euler" if r["integrator"] == "euler" else "semi-implicit")
        env.set_task(*r["xi"])
        env.set_state(np.array([[float(v) for v in r["s0"]]]))
        obs, rew, done, _ = env.step(torch.tensor([r["action"]], dtype=torch.uint8))
        want = np.array([float(v) for v in r["s1"]])
        assert np.max(np.abs(_np(obs)[0] - want)) <= FP64_TOL, r
        assert float(rew[0]) == r["reward"] and bool(done[0]) == r["done"]


@pytest.mark.parametrize("euler", [True, False])
def test_golden_trajectories_fp64_free_running_and_teacher_forced(golden_traj, euler):
    t = golden_traj
    idx = np.where(t["euler"] == euler)[0]
    offs = np.concatenate([[0], np.cumsum(t["length"])])
    L = int(t["length"][idx].max())
    n = len(idx)
    acts = np.zeros((L, n), np.uint8)
    states = np.zeros((L, n, 4))
    dones = np.zeros((L, n), bool)
    valid = np.zeros((L, n), bool)
    for j, e in enumerate(idx):
        T = int(t["length"][e])
        sl = slice(offs[e], offs[e] + T)
        acts[:T, j], states[:T, j], dones[:T, j], valid[:T, j] = t["actions"][sl], t["states"][sl], t["dones"][sl], True
    integ = "euler" if euler else "semi-implicit"
    free, forced = _vec(n, kinematics_integrator=integ), _vec(n, kinematics_integrator=integ)
    for env in (free, forced):
        env.set_task(torch.tensor(t["xi"][idx]))
        env.set_state(t["s0"][idx])
    worst_free = worst_forced = 0.0
    for k in range(L):
        a = torch.tensor(acts[k], device="cuda")
        o1, r1, d1, _ = free.step(a)
        o2, r2, d2, _ = forced.step(a)
        v = valid[k]
        worst_free = max(worst_free, np.max(np.abs(_np(o1)[v] - states[k][v])))
        worst_forced = max(worst_forced, np.max(np.abs(_np(o2)[v] - states[k][v])))
        assert np.array_equal(_np(d1)[v], dones[k][v]) and np.array_equal(_np(d2)[v], dones[k][v]), k
        assert np.all(_np(r1)[v] == 1.0)
        nxt = np.where(v[:, None], states[k], _np(o2))       # teacher forcing: continue from the reference state
        forced.set_state(nxt)
    # per-step (teacher-forced) bound is the 1e-12 of the north star; free-running episodes of up to ~110 steps
    # let the unstable mode amplify a 1-ulp difference by up to ~1e4, hence 1e-10 there
    assert worst_forced <= FP64_TOL and worst_free <= 1e-10, (worst_forced, worst_free)


def test_beyond_done_reward_rule_fp64_and_fp32(golden_dir):
    g = np.load(os.path.join(golden_dir, "cartpole_beyond_done.npz"))
    for dtype, tol in (("float64", FP64_TOL), ("float32", 1e-4)):
        env = _vec(1, dtype=dtype)
        env.set_task(*g["xi"]); env.set_state(g["s0"].reshape(1, 4))
        for k in range(len(g["actions"])):
            obs, rew, done, _ = env.step(torch.tensor([g["actions"][k]], dtype=torch.uint8))
            assert float(rew[0]) == g["rewards"][k] and bool(done[0]) == bool(g["dones"][k]), (dtype, k)
            assert int(env.steps_beyond_done[0]) == g["beyond"][k]
            if k < 12:
                assert np.max(np.abs(_np(obs)[0] - g["states"][k])) <= tol


def test_timelimit_truncation_and_auto_reset(golden_dir):
    """A 500-step survivor.  The GPU env runs its OWN closed loop (a = [w.s > 0] on its own state): replaying the
    reference's actions open-loop would let the pole's unstable mode (e^{4t}) amplify a 1-ulp sin/cos difference
    into a fall long before step 500.  States are compared while that growth is still below 1e-9 (100 steps)."""
    g = np.load(os.path.join(golden_dir, "cartpole_timelimit.npz"))
    env = random_envs.RandomCartPoleVecEnv(1, dtype="float64", seed=11)     # defaults: 500 steps, auto-reset
    env.reset()
    env.set_task(*g["xi"]); env.set_state(g["s0"].reshape(1, 4))
    w = g["w"]
    s = g["s0"]
    for k in range(500):
        a = int(w[0] * s[0] + w[1] * s[1] + w[2] * s[2] + w[3] * s[3] + 0.0 > 0.0)
        obs, rew, done, info = env.step(torch.tensor([a], dtype=torch.uint8))
        s = _np(obs)[0]
        if k < 100:
            assert a == g["actions"][k] and np.max(np.abs(s - g["states"][k])) <= 1e-9, k
        if k < 499:
            assert not bool(done[0]) and not bool(info["TimeLimit.truncated"][0])
            assert int(env.elapsed[0]) == k + 1
    assert bool(done[0]) and bool(info["TimeLimit.truncated"][0]) and float(rew[0]) == 1.0
    assert int(env.elapsed[0]) == 0 and int(env.episode[0]) == 2
    want = c_oracle.init_state(11, 0, 500)           # the obs returned is the RESET obs, keyed by the tick of step 500
    assert np.array_equal(_np(obs)[0], want)


def _random_batch(n, seed):
    rs = np.random.RandomState(seed)
    lo = np.array([2.0, 0.5, 0.05, 0.1]); hi = np.array([20.0, 3.0, 0.3, 1.0])
    xi = lo + (hi - lo) * rs.uniform(size=(n, 4))
    s = rs.uniform(-1, 1, size=(n, 4)) * np.array([2.6, 3.0, 0.23, 3.0])   # straddles both thresholds
    a = rs.randint(0, 2, size=n).astype(np.uint8)
    return xi, s, a


@pytest.mark.parametrize("euler", [True, False])
@pytest.mark.parametrize("n", [1, 2, 3, 5, 31, 1000, 262147])
def test_single_step_vs_c_oracle_all_tail_sizes(n, euler):
    """Vector body + scalar tail (n not a multiple of the 128-bit width), fp64 exact-flag and fp32 tolerance."""
    xi, s, a = _random_batch(n, n)
    ref = np.ascontiguousarray(s.T).copy()
    term = c_oracle.step_batch(ref, np.ascontiguousarray(xi.T), a, euler)
    for dtype, tol in (("float64", FP64_TOL), ("float32", FP32_STEP_TOL)):
        env = _vec(n, dtype=dtype, kinematics_integrator="euler" if euler else "semi")
        env.set_task(torch.tensor(xi)); env.set_state(s)
        obs, rew, done, _ = env.step(torch.tensor(a))
        err = np.abs(_np(obs).astype(np.float64) - ref.T)
        assert err.max() <= tol * max(1.0, np.abs(ref).max()), (dtype, err.max())
        if dtype == "float64":
            assert np.array_equal(_np(done), term)
        else:   # fp32 may flip a flag only when the fp64 state sits within rounding distance of a threshold
            flips = _np(done) != term
            margin = np.minimum(np.abs(np.abs(ref[0]) - 2.4), np.abs(np.abs(ref[2]) - 0.20943951023931953))
            assert np.all(margin[flips] < 1e-5)
        assert np.all(_np(rew) == 1.0) and np.array_equal(_np(env.elapsed), np.ones(n, np.int32))


def test_fp64_closed_loop_with_auto_reset_and_uniform_dr_vs_c_oracle():
    """Free-running: step + TimeLimit + auto-reset + DR resample, 200 steps, every done flag compared."""
    n, K, seed = 20000, 200, 5
    lo = np.array([2.0, 0.5, 0.05, 0.1]); hi = np.array([20.0, 3.0, 0.3, 1.0])
    env = random_envs.RandomCartPoleVecEnv(n, dtype="float64", seed=seed, max_episode_steps=60)
    env.set_dr_distribution("uniform", [2.0, 20.0, 0.5, 3.0, 0.05, 0.3, 0.1, 1.0])
    env.set_dr_training(True)
    obs0 = _np(env.reset()).copy()
    xi0 = _np(env.get_task()).copy()
    # the oracle starts from the same reset: reset() ran at clock tick 0, the K steps run at ticks 1 .. K
    for i in (0, 1, n - 1):
        assert np.array_equal(obs0[i], c_oracle.init_state(seed, i, 0))
        assert np.array_equal(xi0[i], c_oracle.xi_uniform(seed, i, 0, lo, hi))
    st = np.ascontiguousarray(obs0.T); xi = np.ascontiguousarray(xi0.T)
    el = np.zeros(n, np.int32); ep = np.ones(n, np.uint32)
    actions = np.stack([c_oracle.random_actions(n, 0, seed, 1 + k) for k in range(K)])
    out = c_oracle.closed_loop(st, xi, el, ep, seed, 0, 1, K, max_steps=60, actions=actions, lo=lo, hi=hi, log=True)
    worst = 0.0
    for k in range(K):
        a = env.sample_actions()
        assert np.array_equal(_np(a), actions[k])               # action kernel == oracle's Philox bits
        obs, rew, done, info = env.step(a)
        assert np.array_equal(_np(done), out["done"][k]), k
        assert np.array_equal(_np(info["TimeLimit.truncated"]), out["truncated"][k]), k
        worst = max(worst, np.max(np.abs(_np(obs) - out["states"][k].T)))
    assert worst <= 1e-8, worst          # free-running over <= 60-step episodes; resets re-synchronise exactly
    assert np.array_equal(_np(env.elapsed), el) and np.array_equal(_np(env.episode).astype(np.uint32), ep)
    assert np.max(np.abs(_np(env.get_task()) - xi.T)) == 0.0    # uniform DR draws are bit-exact
    assert out["truncated"].sum() > 0 and out["done"].sum() > n  # both ends of the TimeLimit logic were exercised


def test_fp32_teacher_forced_500_steps_and_free_running_growth(golden_dir):
    g = np.load(os.path.join(golden_dir, "cartpole_timelimit.npz"))
    env = _vec(1, dtype="float32")
    env.set_task(*g["xi"])
    prev = g["s0"]
    worst = 0.0
    for k in range(500):
        env.set_state(prev.reshape(1, 4))
        obs, _, _, _ = env.step(torch.tensor([g["actions"][k]], dtype=torch.uint8))
        worst = max(worst, np.max(np.abs(_np(obs)[0].astype(np.float64) - g["states"][k])))
        prev = g["states"][k]
    assert worst <= FP32_STEP_TOL, worst
    # free running from s0 with the reference's action sequence
    env.set_state(g["s0"].reshape(1, 4))
    for k, tol in zip(range(20), [1e-5] * 5 + [1e-4] * 15):
        obs, _, _, _ = env.step(torch.tensor([g["actions"][k]], dtype=torch.uint8))
        assert np.max(np.abs(_np(obs)[0].astype(np.float64) - g["states"][k])) <= tol, k


@pytest.mark.parametrize("euler", [True, False])
def test_fp64_divisions_are_ieee_exact_when_theta_is_zero(euler):
    """With theta = 0, sin = 0 and cos = 1 exactly, so the step is +, *, / only: every bit must equal the oracle's.

    This pins the kernel's division scheme (one correctly rounded reciprocal of total_mass + Markstein's FMA
    correction, renv_cartpole.cuh div_total) to IEEE `/` on 2^20 random (numerator, total_mass) pairs for each of
    the three divisions of random_cartpole.py:183-185."""
    n = 1 << 20
    xi, s, a = _random_batch(n, 77)
    s[:, 2] = 0.0
    s[:, 3] *= 40.0                                  # wide range of numerators
    xi[: n // 2] *= np.random.RandomState(3).uniform(0.01, 100.0, size=(n // 2, 4))   # masses far outside the DR box
    ref = np.ascontiguousarray(s.T).copy()
    term = c_oracle.step_batch(ref, np.ascontiguousarray(xi.T), a, euler)
    env = _vec(n, kinematics_integrator="euler" if euler else "semi")
    env.set_task(torch.tensor(xi)); env.set_state(s)
    obs, _, done, _ = env.step(torch.tensor(a))
    assert np.array_equal(_np(obs), ref.T)
    assert np.array_equal(_np(done), term)


def test_fp64_step_is_bit_identical_to_the_oracle_on_nearly_every_env():
    """sin/cos are the only inexact link: the Taylor kernel agrees with glibc on 99.6 % of arguments (else 1 ulp)."""
    n = 1 << 20
    xi, s, a = _random_batch(n, 78)
    s[:, 2] = np.random.RandomState(4).uniform(-0.2094, 0.2094, n)       # the auto-reset regime
    ref = np.ascontiguousarray(s.T).copy()
    c_oracle.step_batch(ref, np.ascontiguousarray(xi.T), a, True)
    env = _vec(n)
    env.set_task(torch.tensor(xi)); env.set_state(s)
    obs, _, _, _ = env.step(torch.tensor(a))
    got = _np(obs)
    same = np.all(got == ref.T, axis=1)
    assert same.mean() >= 0.98, same.mean()
    assert np.max(np.abs(got - ref.T)) <= 1e-13
