"""Where the time of one scalar step goes: the doorbell round trip alone vs the Python around it."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import random_envs_b200 as renv

env = renv.RandomCartPoleEnv()
env.seed(0); env.reset()
core = env._core
n = 20000
for _ in range(200):
    core._call(1)
t0 = time.perf_counter()
for _ in range(n):
    core._call(0)
print("doorbell round trip (_call, op=step): %.2f us" % (1e6 * (time.perf_counter() - t0) / n))
t0 = time.perf_counter()
for _ in range(n):
    core.step(0)
print("core.step (+ result fetch):            %.2f us" % (1e6 * (time.perf_counter() - t0) / n))
env.reset()
t0 = time.perf_counter()
for i in range(n):
    o, r, d, _ = env.step(i & 1)
    if d:
        env.reset()
print("env.step (bare env, resets incl.):     %.2f us" % (1e6 * (time.perf_counter() - t0) / n))
genv = renv.gym.make("RandomCartPole-v0")
genv.seed(0); genv.reset()
t0 = time.perf_counter()
for i in range(n):
    o, r, d, _ = genv.step(i & 1)
    if d:
        genv.reset()
print("gym.make env.step (TimeLimit wrapper): %.2f us" % (1e6 * (time.perf_counter() - t0) / n))
