"""Time the reference's CPU path on the host cores.  TEST/BENCH INFRASTRUCTURE (see oracle/__init__.py).

What is timed is BASELINE.md section 3 "C2/C3": a gym-0.21 ``SyncVectorEnv``-style serial loop over
``envs_per_proc`` scalar env objects (the Python port of RandomCartPoleEnv under TimeLimit(500)), random
policy, uniform DR over the search bounds resampled on every reset -- replicated in ``procs`` worker
processes (the reference itself has no parallelism; this is the most host throughput its design allows).
The reference source cannot travel to the GPU box, so the port (bit-exact with it, see
tests/test_oracle_cartpole.py) is what runs: ``kind = "port"``.

    python -m oracle.cpu_bench --procs 8 --envs-per-proc 64 --steps 500 --warmup 5
prints one JSON object.
"""
import argparse
import json
import multiprocessing as mp
import os
import platform
import time

import numpy as np

from . import cartpole_port as port, dr_port

SEARCH_LO = np.array([b[0] for b in port.SEARCH_BOUNDS])
SEARCH_HI = np.array([b[1] for b in port.SEARCH_BOUNDS])


def _worker(rank, envs_per_proc, steps, warmup, barrier, out):
    rs = np.random.RandomState(1000 + rank)
    envs = []
    for i in range(envs_per_proc):
        e = port.TimeLimitPort(port.CartPolePort())
        e.env.seed(rank * envs_per_proc + i)
        e.set_task(*dr_port.sample_task("uniform", SEARCH_LO, SEARCH_HI, rng=rs))
        e.reset()
        envs.append(e)

    def resample(env):
        env.set_task(*dr_port.sample_task("uniform", SEARCH_LO, SEARCH_HI, rng=rs))

    for _ in range(warmup):
        port.sync_vector_step(envs, [int(a) for a in rs.randint(0, 2, envs_per_proc)], on_reset=resample)
    barrier.wait()
    t0 = time.perf_counter()
    episodes = 0
    for _ in range(steps):
        _, _, done, _ = port.sync_vector_step(envs, [int(a) for a in rs.randint(0, 2, envs_per_proc)], on_reset=resample)
        episodes += int(done.sum())
    out.put((rank, time.perf_counter() - t0, episodes))


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return platform.processor() or "unknown"


def run(procs=None, envs_per_proc=64, steps=500, warmup=5):
    """Returns dict(value=env-steps/s over all procs, seconds=max worker time, ...)."""
    procs = procs or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(procs)
    out = ctx.Queue()
    workers = [ctx.Process(target=_worker, args=(r, envs_per_proc, steps, warmup, barrier, out)) for r in range(procs)]
    for w in workers:
        w.start()
    results = [out.get() for _ in workers]
    for w in workers:
        w.join()
    seconds = max(r[1] for r in results)
    total = procs * envs_per_proc * steps
    return dict(value=total / seconds, unit="env-steps/s", cores=procs, kind="port", seconds=seconds,
                env_steps=total, episodes=sum(r[2] for r in results), envs_per_proc=envs_per_proc, steps=steps,
                cpu_model=cpu_model(), cpu_count=os.cpu_count(),
                sample="%d procs x %d envs x %d SyncVectorEnv steps, random policy, uniform DR resample on reset"
                       % (procs, envs_per_proc, steps))


def run_for(seconds=10.0, procs=None, envs_per_proc=64):
    """Calibrate on a short run, then time a run of about ``seconds`` wall time."""
    probe = run(procs, envs_per_proc, steps=200, warmup=2)
    per_step = probe["seconds"] / 200
    steps = max(40, int(seconds / per_step))
    return run(procs, envs_per_proc, steps=steps, warmup=2)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=None)
    ap.add_argument("--envs-per-proc", type=int, default=64)
    ap.add_argument("--steps", type=int, default=0)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--seconds", type=float, default=10.0)
    a = ap.parse_args()
    res = run(a.procs, a.envs_per_proc, a.steps, a.warmup) if a.steps > 0 else run_for(a.seconds, a.procs, a.envs_per_proc)
    print(json.dumps(res))
