"""Scalar pure-Python restatement of the RandomCartPole-v0 hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pinned bit-for-bit against the
real reference by ``tests/test_oracle_cartpole.py`` (live, when /root/reference
is mounted) and against ``tests/golden/cartpole_*.npz`` (everywhere).

What is restated and from where (all paths relative to /root/reference):
  constants        random_envs/random_cartpole.py:74-86
  dynamics_step    random_envs/random_cartpole.py:176-205   (D1, D2, D3)
  reward rule      random_envs/random_cartpole.py:207-222   (D4)
  reset            random_envs/random_cartpole.py:226-229   (R1)
  set_task         random_envs/random_cartpole.py:157-166   (T1; polemass_length stays 0.05)
  TimeLimit        gym==0.21.0 wrappers/time_limit.py       [memory; W1]
  sync vector loop gym==0.21.0 vector/sync_vector_env.py    [memory; W2]

The arithmetic is kept in CPython floats with the reference's exact operator
order (including ``** 2`` which CPython routes through libm ``pow``) so that
IEEE results are identical, not merely close.
"""
import math

import numpy as np

# random_cartpole.py:74-86
GRAVITY, CART_MASS, POLE_MASS, POLE_HALF_LENGTH = 9.8, 1.0, 0.1, 0.5
POLEMASS_LENGTH = POLE_MASS * POLE_HALF_LENGTH   # frozen at construction (:79); set_task never refreshes it
FORCE_MAG = 10.0
TAU = 0.02
THETA_THRESHOLD = 12 * 2 * math.pi / 360
X_THRESHOLD = 2.4
MAX_EPISODE_STEPS = 500                          # random_cartpole.py:294
NOMINAL_TASK = (GRAVITY, CART_MASS, POLE_MASS, POLE_HALF_LENGTH)
SEARCH_BOUNDS = ((2.0, 20.0), (0.5, 3.0), (0.05, 0.3), (0.1, 1.0))   # :127-132
LOWER_BOUNDS = (0.1, 0.1, 0.1, 0.1)                                     # :139-145


def dynamics_step(state, xi, action, euler=True):
    """One application of random_cartpole.py:176-205.

    state = (x, x_dot, theta, theta_dot); xi = (gravity, cart_mass, pole_mass,
    pole_length); action in {0, 1}.  Returns (new_state tuple, terminated bool).
    """
    x, x_dot, theta, theta_dot = state
    gravity, cart_mass, pole_mass, pole_length = xi
    total_mass = pole_mass + cart_mass
    push = FORCE_MAG if action == 1 else -FORCE_MAG
    c = math.cos(theta)
    s = math.sin(theta)
    tmp = (push + POLEMASS_LENGTH * theta_dot ** 2 * s) / total_mass
    theta_acc = (gravity * s - c * tmp) / (pole_length * (4.0 / 3.0 - pole_mass * c ** 2 / total_mass))
    x_acc = tmp - POLEMASS_LENGTH * theta_acc * c / total_mass
    if euler:
        x = x + TAU * x_dot
        x_dot = x_dot + TAU * x_acc
        theta = theta + TAU * theta_dot
        theta_dot = theta_dot + TAU * theta_acc
    else:
        x_dot = x_dot + TAU * x_acc
        x = x + TAU * x_dot
        theta_dot = theta_dot + TAU * theta_acc
        theta = theta + TAU * theta_dot
    terminated = bool(x < -X_THRESHOLD or x > X_THRESHOLD
                      or theta < -THETA_THRESHOLD or theta > THETA_THRESHOLD)
    return (x, x_dot, theta, theta_dot), terminated


class CartPolePort:
    """Object-per-env port with the reference's per-step bookkeeping (D4-D6, R1, T1, T2)."""

    task_dim = 4

    def __init__(self, integrator="euler"):
        self.xi = NOMINAL_TASK
        self.kinematics_integrator = integrator
        self.np_random = np.random.RandomState()
        self.state = None
        self.steps_beyond_done = None

    def seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)
        return [seed]

    def set_task(self, *task):
        self.xi = tuple(task[:4])

    def get_task(self):
        return np.array(self.xi)

    def reset(self):
        self.state = tuple(self.np_random.uniform(low=-0.05, high=0.05, size=(4,)))
        self.steps_beyond_done = None
        return np.array(self.state)

    def step(self, action):
        if not (isinstance(action, (int, np.integer)) and 0 <= int(action) < 2):
            raise AssertionError("%r (%s) invalid" % (action, type(action)))
        self.state, terminated = dynamics_step(self.state, self.xi, action,
                                               self.kinematics_integrator == "euler")
        if not terminated:
            reward = 1.0
        elif self.steps_beyond_done is None:
            self.steps_beyond_done = 0
            reward = 1.0
        else:
            self.steps_beyond_done += 1
            reward = 0.0
        return np.array(self.state), reward, terminated, {}


class TimeLimitPort:
    """gym 0.21 TimeLimit semantics [memory]: elapsed>=max => truncated = not done; done = True."""

    def __init__(self, env, max_episode_steps=MAX_EPISODE_STEPS):
        self.env = env
        self.max_episode_steps = max_episode_steps
        self.elapsed = None

    def __getattr__(self, name):
        return getattr(self.env, name)

    def reset(self):
        self.elapsed = 0
        return self.env.reset()

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        self.elapsed += 1
        if self.elapsed >= self.max_episode_steps:
            info["TimeLimit.truncated"] = not done
            done = True
        return obs, reward, done, info


def sync_vector_step(envs, actions, on_reset=None):
    """gym 0.21 SyncVectorEnv.step_wait [memory]: serial loop, auto-reset on done.

    ``on_reset(env)`` is the harness hook that resamples xi before ``reset`` --
    the reference CartPole never does it itself (random_cartpole.py:226-229),
    while the README (README.md:9) and every MuJoCo env do.
    """
    n = len(envs)
    obs = np.empty((n, 4))
    rew = np.empty(n)
    done = np.zeros(n, dtype=bool)
    trunc = np.zeros(n, dtype=bool)
    for i in range(n):
        o, rew[i], done[i], info = envs[i].step(actions[i])
        trunc[i] = info.get("TimeLimit.truncated", False)
        if done[i]:
            if on_reset is not None:
                on_reset(envs[i])
            o = envs[i].reset()
        obs[i] = o
    return obs, rew, done, trunc
