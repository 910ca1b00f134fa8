"""Time the fused rollout (2^24 envs x 500 steps): python profiles/time_rollout.py [float32|float64]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import random_envs_b200 as renv  # noqa: E402

dtypes = sys.argv[1:] or ["float32", "float64"]
for dtype in dtypes:
    for w in ((0.1, 0.1, 1.0, 0.3), (0.0, 0.0, 1.0, 0.0), None):
        env = renv.RandomCartPoleVecEnv(1 << 24, dtype=dtype, seed=2)
        env.set_dr_distribution("uniform", [2.0, 20.0, 0.5, 3.0, 0.05, 0.3, 0.1, 1.0]); env.set_dr_training(True); env.reset()
        env.rollout(w, 0.0, 10); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); env.rollout(w, 0.0, 500); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("%s w=%s: %.2f ms  %.3e env-steps/s" % (dtype, w, ms, (1 << 24) * 500 / ms * 1e3))
        del env
