"""Tiny driver for ncu captures: runs ONE kernel family a few times at a chosen size.

    python profiles/drive.py step --n 16777216 --dtype float32 --iters 6
    python profiles/drive.py rollout --n 16777216 --dtype float32 --iters 2 --K 500
    python profiles/drive.py sample --n 16777216 --dr truncnorm --iters 3
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import random_envs_b200 as renv  # noqa: E402

SEARCH = [2.0, 20.0, 0.5, 3.0, 0.05, 0.3, 0.1, 1.0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("kernel", choices=["step", "rollout", "sample"])
    ap.add_argument("--n", type=int, default=1 << 24)
    ap.add_argument("--dtype", default="float32")
    ap.add_argument("--iters", type=int, default=6)
    ap.add_argument("--K", type=int, default=500)
    ap.add_argument("--dr", default="uniform")
    ap.add_argument("--policy", default="survive", choices=["survive", "resetheavy", "random"])
    ap.add_argument("--warm", type=int, default=30, help="env-steps before the measured launches (mix of episode ages)")
    ap.add_argument("--lean", action="store_true", help="the 54-byte step (uint16 TimeLimit counter, no reward store)")
    ap.add_argument("--tile-ordering", default="auto", choices=["auto", "on", "off"])
    a = ap.parse_args()
    if a.kernel == "sample":
        s = renv.TaskSampler("RandomHumanoid-v0")
        distr = []
        for v in renv.HUMANOID_NOMINAL:
            distr += [0.5 * v, 1.5 * v] if a.dr == "uniform" else [v, 0.1 * v]
        s.set_dr_distribution(a.dr, distr)
        buf = torch.empty((a.n, 30), dtype=getattr(torch, a.dtype), device="cuda")
        for _ in range(a.iters):
            s.sample_tasks_tensor(a.n, out=buf)
        torch.cuda.synchronize()
        return
    tile = {"auto": "auto", "on": True, "off": False}[a.tile_ordering]
    env = renv.RandomCartPoleVecEnv(a.n, dtype=a.dtype, seed=0, track_truncated=False, track_episodes=False, lean=a.lean,
                                    tile_ordering=tile)
    env.set_dr_distribution("uniform", SEARCH); env.set_dr_training(True); env.reset()
    if a.kernel == "step":
        act = env.sample_actions().clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(a.warm):
            env.step(act)
        torch.cuda.synchronize(); e0.record()
        for _ in range(a.iters):
            env.step(act)
        e1.record(); torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / a.iters
        b = (54 if a.lean else 62) if a.dtype == "float32" else 114
        print("step %s n=%d: %.1f us/launch, %.0f GB/s algorithmic" % (a.dtype, a.n, us, b * a.n / us / 1e3))
    else:
        w = {"survive": (0.1, 0.1, 1.0, 0.3), "resetheavy": (0.0, 0.0, 1.0, 0.0), "random": None}[a.policy]
        for _ in range(a.iters):
            env.rollout(w, 0.0, a.K)
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
