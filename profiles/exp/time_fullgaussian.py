"""Time the 30-dim fullgaussian sampler (tensor-core path vs RENV_FULLGAUSS_TENSOR=0 CUDA-core path)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import random_envs_b200 as renv  # noqa: E402

s = renv.TaskSampler("RandomHumanoid-v0")
a = np.random.RandomState(0).randn(30, 30) * 0.1
s.set_dr_distribution("fullgaussian", {"mean": np.full(30, 2.0), "cov": a @ a.T + 0.05 * np.eye(30)})
n = 1 << 24
buf = torch.empty((n, 30), dtype=torch.float32, device="cuda")
for knob in ("1", "0"):
    os.environ["RENV_FULLGAUSS_TENSOR"] = knob
    s.sample_tasks_tensor(n, out=buf); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        s.sample_tasks_tensor(n, out=buf)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("fullgaussian %s 2^24 x 30: %.3f ms  %.0f GB/s  %.3e xi/s" % ("tcgen05" if knob == "1" else "fma-chain", ms, n * 120 / ms / 1e6, n / ms * 1e3))
    s.check_dr_violations()
