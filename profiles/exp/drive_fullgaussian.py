import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import random_envs_b200 as renv
s = renv.TaskSampler("RandomHumanoid-v0")
a = np.random.RandomState(0).randn(30, 30) * 0.1
s.set_dr_distribution("fullgaussian", {"mean": np.full(30, 2.0), "cov": a @ a.T + 0.05 * np.eye(30)})
n = 1 << 22
buf = torch.empty((n, 30), dtype=torch.float32, device="cuda")
for _ in range(3):
    s.sample_tasks_tensor(n, out=buf)
torch.cuda.synchronize()
s.check_dr_violations()
