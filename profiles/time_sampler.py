"""Time the DR sampler kernel: 2^24 x 30 (humanoid) fp32 for each dr_type."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import random_envs_b200 as renv  # noqa: E402

n = 1 << 24
for dtype in (torch.float32, torch.float64):
    for d in ("uniform", "gaussian", "truncnorm"):
        s = renv.TaskSampler("RandomHumanoid-v0")
        distr = []
        for v in renv.HUMANOID_NOMINAL:
            distr += [0.5 * v, 1.5 * v] if d == "uniform" else [v, 0.1 * v]
        s.set_dr_distribution(d, distr)
        nn = n if dtype == torch.float32 else n // 2
        buf = torch.empty((nn, 30), dtype=dtype, device="cuda")
        s.sample_tasks_tensor(nn, out=buf); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            s.sample_tasks_tensor(nn, out=buf)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print("%s %-9s %.3f ms  %.0f GB/s  %.3e xi/s" % (str(dtype)[6:], d, ms, buf.numel() * buf.element_size() / ms / 1e6, nn / ms * 1e3))
        del buf
