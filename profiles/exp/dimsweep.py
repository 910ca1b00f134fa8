import sys, os
sys.path.insert(0, os.getcwd())
import torch
from random_envs_b200 import _lib, _device
import random_envs_b200 as renv
renv.load_library()
dev = torch.device("cuda", 0)
n = 1 << 24
for dim in (4, 8, 13, 16, 23, 28, 30, 32):
    for dr in ("uniform", "truncnorm"):
        a = [1.0] * dim; b = [2.0] * dim if dr == "uniform" else [0.1] * dim
        cfg = _lib.make_dr_cfg(dr, a, b, [0.1] * dim)
        out = torch.empty((n, dim), dtype=torch.float32, device=dev)
        viol = torch.zeros(1, dtype=torch.int64, device=dev)
        f = lambda: _lib.call("renv_dr_sample_f32", _device.ptr(out), n, cfg, 1, 0, 0, _device.ptr(viol), _device.stream_ptr(dev))
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        blocks = n * ((dim + 3) // 4)
        print("dim %2d %-9s %.3f ms  %.0f GB/s  %.3e philox-blocks/s" % (dim, dr, ms, n * dim * 4 / ms / 1e6, blocks / ms * 1e3))
        del out
