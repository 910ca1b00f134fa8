"""Generate tests/golden/* by executing the REAL reference source.  Build container only.

    python -m oracle.make_golden            # rewrites tests/golden/

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference ships no golden vectors
(SURVEY.md section 4), so these are produced by running its unmodified
``RandomCartPoleEnv.step`` / ``RandomEnv.sample_task`` (loaded by oracle/reference_loader.py)
on seeded inputs.  The fixtures are small, committed, and are what the oracle restatements and
the CUDA kernels are checked against on machines that have no /root/reference.

Files written
  cartpole_known_answers.json  the 8 one-step rows of SURVEY.md section 8c (repr precision)
  cartpole_traj.npz            free-running random-policy episodes, uniform DR over the search
                               bounds, both integrators: xi, s0, actions, every state, reward, done
  cartpole_beyond_done.npz     one env stepped 40 times past termination (reward 1,1,...,1,0,0 rule)
  cartpole_timelimit.npz       a 500-step survivor (bang-bang on theta) for the TimeLimit edge
  sampler_control_flow.json    scripted-draw known answers for the truncnorm / gaussian retry loops
  sampler_reference_draws.npz  iid draws of the reference's own sample_task (uniform, gaussian,
                               truncnorm incl. a lower-bound point-mass case, fullgaussian with clipping)
                               for two-sample KS
  cartpole_episode_lengths.npz the reference's own demo loop (test_random_policy.py:25-32: action_space.sample(),
                               reset on done) under uniform DR over the search bounds, resampled at every reset through
                               the reference's set_random_task: 20 000 episode lengths (both integrators) -- the
                               end-to-end law of dynamics + termination + reset + DR, for two-sample KS against the GPU
"""
import json
import os

import numpy as np

from . import reference_loader as rl

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SEARCH = [2.0, 20.0, 0.5, 3.0, 0.05, 0.3, 0.1, 1.0]   # random_cartpole.py:127-132, interleaved


def _inject(env, s0, xi):
    env.state = np.array(s0, dtype=np.float64).copy()
    env.steps_beyond_done = None
    env.set_task(*[float(v) for v in xi])


def known_answers():
    s0 = np.random.RandomState(0).uniform(-0.05, 0.05, 4)
    rows = []
    for xi in [(9.8, 1.0, 0.1, 0.5), (5.0, 2.0, 0.2, 0.8)]:
        for integ in ("euler", "semi-implicit"):
            for a in (1, 0):
                env = rl.make_cartpole()
                env.kinematics_integrator = integ
                _inject(env, s0, xi)
                obs, rew, done, _ = env.step(a)
                rows.append(dict(xi=list(xi), integrator=integ, action=a, s0=[repr(float(v)) for v in s0],
                                 s1=[repr(float(v)) for v in obs], reward=rew, done=bool(done),
                                 polemass_length=env.polemass_length))
    return rows


def trajectories(n_episodes=160, seed=1234):
    rs = np.random.RandomState(seed)
    xs, s0s, acts, states, rewards, dones, lens, integ = [], [], [], [], [], [], [], []
    lo, hi = np.array(SEARCH[0::2]), np.array(SEARCH[1::2])
    for e in range(n_episodes):
        env = rl.make_cartpole()
        euler = e % 2 == 0
        env.kinematics_integrator = "euler" if euler else "semi-implicit-euler"
        xi = lo + (hi - lo) * rs.uniform(size=4)
        s0 = rs.uniform(-0.05, 0.05, 4)
        _inject(env, s0, xi)
        T = 0
        ep_a, ep_s, ep_r, ep_d = [], [], [], []
        while True:
            a = int(rs.randint(2))
            obs, r, d, _ = env.step(a)
            ep_a.append(a); ep_s.append(obs); ep_r.append(r); ep_d.append(d)
            T += 1
            if d or T >= 500:
                break
        xs.append(xi); s0s.append(s0); lens.append(T); integ.append(euler)
        acts.append(np.array(ep_a, np.uint8)); states.append(np.array(ep_s)); rewards.append(np.array(ep_r))
        dones.append(np.array(ep_d, bool))
    return dict(xi=np.array(xs), s0=np.array(s0s), length=np.array(lens, np.int32), euler=np.array(integ, bool),
                actions=np.concatenate(acts), states=np.concatenate(states), rewards=np.concatenate(rewards),
                dones=np.concatenate(dones))


def beyond_done():
    """Stepping past termination: reward 1 (first done), then 0 while done, and -- because the state keeps
    integrating -- done=False / reward=1 again when the cart drifts back inside the track (:207-222)."""
    s0 = np.array([2.39, 0.5, -0.05, 0.0])
    xi = np.array([9.8, 1.0, 0.1, 0.5])
    actions = np.zeros(40, np.uint8)
    env = rl.make_cartpole()
    _inject(env, s0, xi)
    S, R, D, B = [], [], [], []
    for k in range(40):
        obs, r, d, _ = env.step(int(actions[k]))
        S.append(obs); R.append(r); D.append(d)
        B.append(-1 if env.steps_beyond_done is None else env.steps_beyond_done)
    D = np.array(D, bool)
    first = int(np.argmax(D))
    assert D.any() and (~D[first:]).any() and D[-1], "scenario must leave, re-enter and leave the thresholds"
    return dict(s0=s0, xi=xi, actions=actions, states=np.array(S), rewards=np.array(R), dones=D,
                beyond=np.array(B, np.int32))


def timelimit_survivor():
    """Closed loop a=[theta + 0.3*theta_dot > 0 ...] long enough to hit 500 steps without terminating."""
    env = rl.make_cartpole()
    w = np.array([0.1, 0.1, 1.0, 0.3])
    s0 = np.array([0.01, -0.02, 0.03, 0.01])
    _inject(env, s0, (9.8, 1.0, 0.1, 0.5))
    A, S, D = [], [], []
    s = s0
    for k in range(520):
        a = int(w[0] * s[0] + w[1] * s[1] + w[2] * s[2] + w[3] * s[3] + 0.0 > 0.0)
        s, r, d, _ = env.step(a)
        A.append(a); S.append(s); D.append(d)
    assert not any(D), "survivor policy failed"
    return dict(s0=s0, xi=np.array([9.8, 1.0, 0.1, 0.5]), w=w, actions=np.array(A, np.uint8), states=np.array(S),
                dones=np.array(D, bool))


class _Scripted:
    """Deterministic stand-ins for truncnorm.rvs / np.random.randn that count consumption."""

    def __init__(self, values):
        self.values = list(values)
        self.used = 0

    def rvs(self, a, b, loc=0.0, scale=1.0):
        v = self.values[self.used]
        self.used += 1
        return v

    def randn(self):
        v = self.values[self.used]
        self.used += 1
        return v


def sampler_control_flow():
    re_mod, _ = rl.load()
    out = []
    # truncnorm: one dim, lb = 0.1
    for script in ([-1.0, -1.0, -1.0, 5.0], [-1.0, -1.0, 3.0, 9.0], [-1.0, 2.0, 9.0, 9.0], [0.5, 9.0, 9.0, 9.0],
                   [0.1, 9.0, 9.0, 9.0], [0.09999, 0.09999, 0.09999, 0.09999]):
        env = rl.make_sampler_env(1, [0.1])
        env.set_dr_distribution("truncnorm", [1.0, 0.5])
        fake = _Scripted(script)
        saved = re_mod.truncnorm
        re_mod.truncnorm = fake
        try:
            val = float(env.sample_task()[0])
        finally:
            re_mod.truncnorm = saved
        out.append(dict(kind="truncnorm", lb=0.1, script=script, result=val, consumed=fake.used))
    # gaussian: mean 0, std 1 so the scripted randn IS the candidate value
    for script in ([-5.0, -5.0, -5.0, 7.0], [-5.0, -5.0, 7.0, 9.0], [-5.0, 0.1, 9.0, 9.0], [0.2, 9.0, 9.0, 9.0]):
        env = rl.make_sampler_env(1, [0.1])
        env.set_dr_distribution("gaussian", [0.0, 1.0])
        fake = _Scripted(script)
        saved = re_mod.np.random.randn
        re_mod.np.random.randn = fake.randn
        try:
            try:
                val, err = float(env.sample_task()[0]), None
            except Exception as exc:  # noqa: BLE001 - the reference raises a bare Exception
                val, err = None, str(exc)
        finally:
            re_mod.np.random.randn = saved
        out.append(dict(kind="gaussian", script=script, result=val, error=err, consumed=fake.used))
    return out


def sampler_reference_draws(n=4000):
    np.random.seed(20261018)
    out = {}
    env = rl.make_cartpole()
    env.set_dr_distribution("uniform", SEARCH)
    out["uniform_params"] = np.array(SEARCH)
    out["uniform"] = env.sample_tasks(n)
    g = [9.8, 0.98, 1.0, 0.1, 0.2, 0.05, 0.5, 0.05]   # pole_mass: floor 0.1 is 2 sigma below the mean
    env.set_dr_distribution("gaussian", g)
    out["gaussian_params"] = np.array(g)
    out["gaussian"] = env.sample_tasks(n)
    t = [9.8, 0.98, 1.0, 0.1, 0.1, 0.02, 0.5, 0.05]   # pole_mass mean == lb=0.1 -> p = 1/2, point mass 1/8 at lb
    env.set_dr_distribution("truncnorm", t)
    out["truncnorm_params"] = np.array(t)
    out["truncnorm_lb"] = np.array([env.get_task_lower_bound(i) for i in range(4)])
    out["truncnorm"] = env.sample_tasks(n)
    # fullgaussian: normalised-space mean/cov (random_env.py:123-127,192-198); dim 3 has a wide variance so that
    # the clip to [0, 4] is exercised on both sides
    mean = np.array([2.0, 1.0, 3.0, 2.0])
    a = np.array([[0.30, 0.00, 0.00, 0.00], [0.10, 0.25, 0.00, 0.00], [-0.05, 0.08, 0.20, 0.00], [0.40, -0.30, 0.20, 1.50]])
    cov = a @ a.T
    env.set_dr_distribution("fullgaussian", {"mean": mean, "cov": cov})
    out["fullgaussian_mean"], out["fullgaussian_cov"] = mean, cov
    out["fullgaussian"] = env.sample_tasks(n)
    return out


def episode_lengths(n_episodes=20000, seed=2024):
    """Unmodified reference classes end to end: env.set_random_task() (its own sample_task on numpy's global state),
    env.reset() (its own np_random), env.step(action_space.sample()) until done; TimeLimit(500) as gym.make adds."""
    out = {}
    for integ in ("euler", "semi-implicit-euler"):
        env = rl.make_cartpole()
        env.kinematics_integrator = integ
        env.seed(seed)
        env.action_space.seed(seed)
        np.random.seed(seed)
        env.set_dr_distribution("uniform", SEARCH)
        env.set_dr_training(True)
        lens = np.zeros(n_episodes, np.int16)
        for e in range(n_episodes):
            env.set_random_task()          # README.md:9 behaviour; CartPole's own reset forgets it
            env.reset()
            T = 0
            while True:
                _, _, d, _ = env.step(env.action_space.sample())
                T += 1
                if d or T >= 500:
                    break
            lens[e] = T
        out["euler" if integ == "euler" else "semi_implicit"] = lens
    return out


def main():
    assert rl.available(), "reference not mounted"
    os.makedirs(GOLDEN, exist_ok=True)
    with open(os.path.join(GOLDEN, "cartpole_known_answers.json"), "w") as f:
        json.dump(known_answers(), f, indent=1)
    np.savez_compressed(os.path.join(GOLDEN, "cartpole_traj.npz"), **trajectories())
    np.savez_compressed(os.path.join(GOLDEN, "cartpole_beyond_done.npz"), **beyond_done())
    np.savez_compressed(os.path.join(GOLDEN, "cartpole_timelimit.npz"), **timelimit_survivor())
    with open(os.path.join(GOLDEN, "sampler_control_flow.json"), "w") as f:
        json.dump(sampler_control_flow(), f, indent=1)
    np.savez_compressed(os.path.join(GOLDEN, "sampler_reference_draws.npz"), **sampler_reference_draws())
    np.savez_compressed(os.path.join(GOLDEN, "cartpole_episode_lengths.npz"), **episode_lengths())
    for name in sorted(os.listdir(GOLDEN)):
        print(name, os.path.getsize(os.path.join(GOLDEN, name)))


if __name__ == "__main__":
    main()
