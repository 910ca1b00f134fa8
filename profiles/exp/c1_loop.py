"""The reference's demo loop (test_random_policy.py:25-32) on the drop-in gym env: us per step, look-ahead on / off
(RENV_SCALAR_LOOKAHEAD=0)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import random_envs_b200 as renv

env = renv.gym.make("RandomCartPole-v0")
env.seed(0); np.random.seed(0)
env.set_dr_distribution("uniform", [2.0, 20.0, 0.5, 3.0, 0.05, 0.3, 0.1, 1.0]); env.set_dr_training(True)
env.reset()
for rep in range(3):
    n, episodes = 10000, 0
    t0 = time.perf_counter()
    for i in range(n):
        o, r, d, _ = env.step(env.action_space.sample())
        if d:
            env.reset(); episodes += 1
    dt = time.perf_counter() - t0
    print("lookahead=%s  %.2f us/step  (%d episodes)" % (os.environ.get("RENV_SCALAR_LOOKAHEAD", "1"), 1e6 * dt / n, episodes))
