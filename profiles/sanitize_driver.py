"""Small driver that touches every kernel at awkward sizes; run under compute-sanitizer (one tool per gpurun call):

    compute-sanitizer --tool memcheck python profiles/sanitize_driver.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import random_envs_b200 as renv  # noqa: E402

SEARCH = [2.0, 20.0, 0.5, 3.0, 0.05, 0.3, 0.1, 1.0]
for dtype in ("float32", "float64"):
    for n in (1, 2, 3, 5, 255, 257, 1031):
        for noisy in (False, True):
            for dr in (("uniform", SEARCH), ("truncnorm", [9.8, 1.0, 1.0, 0.5, 0.1, 0.05, 0.5, 0.2]),
                       ("gaussian", [9.8, 1.0, 1.0, 0.1, 0.3, 0.02, 0.5, 0.05])):
                env = renv.RandomCartPoleVecEnv(n, dtype=dtype, seed=n, max_episode_steps=9, noisy=noisy)
                env.set_dr_distribution(*dr); env.set_dr_training(True)
                env.reset()
                for _ in range(12):
                    env.step(env.sample_actions())
                env.reset(mask=torch.arange(n, device="cuda") % 2)
                if not noisy:
                    env.rollout((0.0, 0.0, 1.0, 0.0), 0.0, 37)
                    env.rollout((0.1, 0.1, 1.0, 0.3), 0.0, 3)
                env2 = renv.RandomCartPoleVecEnv(n, dtype=dtype, seed=n, auto_reset=False, max_episode_steps=0, noisy=noisy)
                env2.reset()
                for _ in range(5):
                    env2.step(torch.ones(n, dtype=torch.uint8, device="cuda"))
for env_id in sorted(renv.XI_TABLES):
    for dr_type in ("uniform", "truncnorm", "gaussian"):
        t = renv.XI_TABLES[env_id]
        lo = [b[0] for b in t.search_bounds]; hi = [b[1] for b in t.search_bounds]
        s = renv.TaskSampler(env_id)
        distr = []
        for a, b in zip(lo, hi):
            distr += [a, b] if dr_type == "uniform" else [(a + b) / 2, (b - a) / 10]
        s.set_dr_distribution(dr_type, distr)
        for n in (1, 7, 129, 2049, 4099):
            for dt in (torch.float32, torch.float64):
                s.sample_tasks_tensor(n, dtype=dt)
s = renv.TaskSampler("RandomHumanoid-v0")
import numpy as np  # noqa: E402
s.set_dr_distribution("fullgaussian", {"mean": np.full(30, 2.0), "cov": np.eye(30) * 0.25})
for n in (1, 63, 65, 1000):
    for dt in (torch.float32, torch.float64):
        s.sample_tasks_tensor(n, dtype=dt)
# round 2: random-policy rollout across action-block boundaries, lean / tile-ordered step, flag bits on the host path,
# the resident scalar env with and without look-ahead
for n in (1, 33, 1031, 3 * 1024 + 5):
    env = renv.RandomCartPoleVecEnv(n, seed=n, max_episode_steps=11)
    env.set_dr_distribution("uniform", SEARCH); env.set_dr_training(True); env.reset()
    for _ in range(125):
        env.step(env.sample_actions())
    env.rollout(None, 0.0, 140)
    env.step_host(np.zeros(n, dtype=np.uint8)); env.step_host(np.ones(n, dtype=np.uint8))
    lean = renv.RandomCartPoleVecEnv(n, seed=n, lean=True, tile_ordering=True)
    lean.set_dr_distribution("uniform", SEARCH); lean.set_dr_training(True); lean.reset()
    for _ in range(40):
        lean.step(lean.sample_actions())
for look in ("1", "0"):
    os.environ["RENV_SCALAR_LOOKAHEAD"] = look
    for noisy in (False, True):
        e = renv.RandomCartPoleEnv(noisy=noisy)
        e.seed(1); e.set_dr_distribution("uniform", SEARCH); e.set_dr_training(True); e.reset()
        for k in range(300):
            _, _, d, _ = e.step(k & 1)
            if d:
                e.reset()
        e.close()
torch.cuda.synchronize()
print("sanitize driver done")
