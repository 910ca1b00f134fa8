"""Time the reference's CPU path on the host cores.  TEST/BENCH INFRASTRUCTURE (see oracle/__init__.py).

What is timed:

  C1  one RandomCartPole-v0 env under TimeLimit(500), uniform DR over the search bounds resampled on every reset,
      random policy, 10,000 steps -- the loop of the reference's test_random_policy.py:25-32 (BASELINE.json configs[0]);
  C2  a gym-0.21 ``SyncVectorEnv``-style serial loop over ``envs_per_proc`` scalar env objects, same DR and policy;
  C3  C2 replicated in ``procs`` worker processes (the reference has no parallelism of its own; this is the most host
      throughput its design allows) -- the figure bench.py quotes as ``cpu_baseline`` and as ``--impl reference``.

The env objects are the REFERENCE's own ``RandomCartPoleEnv`` (``kind = "reference"``: ``random_envs/random_env.py`` and
``random_cartpole.py`` exec'd unmodified from /root/reference or from the copies ``oracle/make_ref.py`` put under
``oracle/_ref/``) whenever those sources are present, else the bit-exact Python port (``kind = "port"``).  gym 0.21 is
not in the image: ``TimeLimit`` and the vector loop are the restatements in ``cartpole_port.py`` either way, and the
harness calls ``env.set_random_task()`` around ``reset()`` because the reference CartPole's own ``reset`` forgets to
(random_cartpole.py:226-229 vs README.md:9).

    python -m oracle.cpu_bench --seconds 10          # C3 for about 10 s, prints one JSON object
    python -m oracle.cpu_bench --c1                  # C1
"""
import argparse
import json
import multiprocessing as mp
import os
import platform
import time

import numpy as np

from . import cartpole_port as port, dr_port, reference_loader

SEARCH_LO = np.array([b[0] for b in port.SEARCH_BOUNDS])
SEARCH_HI = np.array([b[1] for b in port.SEARCH_BOUNDS])
SEARCH_INTERLEAVED = [v for b in port.SEARCH_BOUNDS for v in b]      # set_dr_distribution('uniform', [lo0, hi0, ...])


def kind():
    return "reference" if reference_loader.available() else "port"


def make_env(seed, rs):
    """One env under TimeLimit(500) with uniform DR over the search bounds; returns (env, resample_fn)."""
    if reference_loader.available():
        base = reference_loader.make_cartpole()                 # the reference's RandomCartPoleEnv, unmodified
        base.set_dr_distribution("uniform", SEARCH_INTERLEAVED)  # random_env.py:72-90
        base.set_dr_training(True)
        env = port.TimeLimitPort(base)
        base.seed(seed)

        def resample(e):
            e.env.set_random_task()                              # random_env.py:37-39 -> sample_task :148-151 (global np.random)
        resample(env)
        return env, resample
    env = port.TimeLimitPort(port.CartPolePort())
    env.env.seed(seed)

    def resample(e):
        e.set_task(*dr_port.sample_task("uniform", SEARCH_LO, SEARCH_HI, rng=rs))
    resample(env)
    return env, resample


def run_c1(steps=10000, seed=0):
    """BASELINE.json configs[0]: single env, random policy, `steps` steps (test_random_policy.py:25-32)."""
    np.random.seed(seed)
    rs = np.random.RandomState(seed)
    env, resample = make_env(seed, rs)
    env.reset()
    episodes = 0
    t0 = time.perf_counter()
    for _ in range(steps):
        _, _, done, _ = env.step(int(rs.randint(0, 2)))          # action_space.sample()
        if done:
            resample(env)
            env.reset()
            episodes += 1
    dt = time.perf_counter() - t0
    return dict(value=steps / dt, unit="env-steps/s", us_per_step=1e6 * dt / steps, steps=steps, episodes=episodes,
                mean_episode_length=steps / max(1, episodes), seconds=dt, cores=1, kind=kind(), cpu_model=cpu_model(),
                sample="1 env x %d steps, random policy, uniform DR resample on reset (test_random_policy.py loop)" % steps)


def _worker(rank, envs_per_proc, steps, warmup, barrier, out, seconds=None):
    rs = np.random.RandomState(1000 + rank)
    np.random.seed(1000 + rank)             # the reference draws xi from the global numpy state
    envs, resample = [], None
    for i in range(envs_per_proc):
        e, resample = make_env(rank * envs_per_proc + i, rs)
        e.reset()
        envs.append(e)

    for _ in range(warmup):
        port.sync_vector_step(envs, [int(a) for a in rs.randint(0, 2, envs_per_proc)], on_reset=resample)
    barrier.wait()
    t0 = time.perf_counter()
    episodes, done_steps = 0, 0
    while done_steps < steps if seconds is None else time.perf_counter() - t0 < seconds:
        _, _, done, _ = port.sync_vector_step(envs, [int(a) for a in rs.randint(0, 2, envs_per_proc)], on_reset=resample)
        episodes += int(done.sum())
        done_steps += 1
    out.put((rank, time.perf_counter() - t0, episodes, done_steps))


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return platform.processor() or "unknown"


def run(procs=None, envs_per_proc=64, steps=500, warmup=5, seconds=None):
    """``steps`` SyncVectorEnv steps per worker -- or, with ``seconds``, as many as fit into that wall time.
    Returns dict(value=env-steps/s over all procs, seconds=max worker time, ...)."""
    procs = procs or os.cpu_count() or 1
    if reference_loader.available():
        reference_loader.load()             # exec the reference modules once, before the workers are forked
    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(procs)
    out = ctx.Queue()
    workers = [ctx.Process(target=_worker, args=(r, envs_per_proc, steps, warmup, barrier, out, seconds))
               for r in range(procs)]
    for w in workers:
        w.start()
    results = [out.get() for _ in workers]
    for w in workers:
        w.join()
    elapsed = max(r[1] for r in results)
    total = envs_per_proc * sum(r[3] for r in results)
    mean_steps = sum(r[3] for r in results) / procs
    return dict(value=total / elapsed, unit="env-steps/s", cores=procs, kind=kind(), seconds=elapsed,
                env_steps=total, episodes=sum(r[2] for r in results), envs_per_proc=envs_per_proc, steps=mean_steps,
                cpu_model=cpu_model(), cpu_count=os.cpu_count(),
                sample="%d procs x %d envs x %.0f SyncVectorEnv steps (%.1f s), random policy, uniform DR resample on reset"
                       % (procs, envs_per_proc, mean_steps, elapsed))


def run_for(seconds=10.0, procs=None, envs_per_proc=64):
    """Every worker steps its envs for ``seconds`` of wall time (after 20 warm-up steps)."""
    return run(procs, envs_per_proc, steps=0, warmup=20, seconds=seconds)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=None)
    ap.add_argument("--envs-per-proc", type=int, default=64)
    ap.add_argument("--steps", type=int, default=0)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--c1", action="store_true")
    a = ap.parse_args()
    if a.c1:
        res = run_c1()
    else:
        res = run(a.procs, a.envs_per_proc, a.steps, a.warmup) if a.steps > 0 else run_for(a.seconds, a.procs, a.envs_per_proc)
    print(json.dumps(res))
