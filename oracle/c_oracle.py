"""Build recipe + ctypes bindings for ``cartpole_oracle.c``.  TEST INFRASTRUCTURE ONLY.

``build()`` compiles ``oracle/cartpole_oracle.c`` into ``oracle/_build/liboracle.so`` with
gcc.  Flags matter for bit-exactness with CPython: ``-ffp-contract=off`` (no FMA fusion),
``-fno-builtin`` (keep ``pow(x, 2.0)`` a libm call as CPython's ``x ** 2`` is), no
``-ffast-math``.  The .so is git-ignored but travels to the GPU box with the snapshot.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "cartpole_oracle.c")
OUT_DIR = os.path.join(_HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle.so")

PURPOSE_INIT, PURPOSE_XI, PURPOSE_ACTION, PURPOSE_TASKS, PURPOSE_OBS = 0, 1, 2, 3, 4


def build(force=False):
    if not force and os.path.isfile(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-builtin",
           "-fno-fast-math", "-o", LIB, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t)) if a is not None else None


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def step_batch(state, xi, action, euler=True):
    """state, xi: (4, n) float64 (state is updated IN PLACE); action (n,) uint8 -> terminated (n,) bool."""
    assert state.dtype == np.float64 and state.flags.c_contiguous and state.shape[0] == 4
    xi = _f64(xi)
    action = np.ascontiguousarray(action, dtype=np.uint8)
    n = state.shape[1]
    term = np.zeros(n, dtype=np.uint8)
    lib().oracle_step_batch(ctypes.c_int64(n), _p(state, ctypes.c_double), _p(xi, ctypes.c_double),
                            _p(action, ctypes.c_uint8), ctypes.c_int(1 if euler else 0), _p(term, ctypes.c_uint8))
    return term.astype(bool)


def philox(ctr, key):
    c = (ctypes.c_uint32 * 4)(*ctr)
    k = (ctypes.c_uint32 * 2)(*key)
    o = (ctypes.c_uint32 * 4)()
    lib().oracle_philox4x32_10(c, k, o)
    return tuple(int(v) for v in o)


def init_state(seed, env_id, episode, dtype=np.float64):
    if np.dtype(dtype) == np.float64:
        o = (ctypes.c_double * 4)()
        lib().oracle_init_state_f64(ctypes.c_uint64(seed), ctypes.c_uint64(env_id), ctypes.c_uint64(episode), o)
    else:
        o = (ctypes.c_float * 4)()
        lib().oracle_init_state_f32(ctypes.c_uint64(seed), ctypes.c_uint64(env_id), ctypes.c_uint64(episode), o)
    return np.array(list(o), dtype=dtype)


def xi_uniform(seed, sample_id, episode, lo, hi, purpose=PURPOSE_XI, dtype=np.float64):
    lo, hi = _f64(lo), _f64(hi)
    dim = lo.shape[0]
    out = np.zeros(dim, dtype=dtype)
    if np.dtype(dtype) == np.float64:
        f, t = lib().oracle_xi_uniform_f64, ctypes.c_double
    else:
        f, t = lib().oracle_xi_uniform_f32, ctypes.c_float
    f(ctypes.c_uint64(seed), ctypes.c_uint64(sample_id), ctypes.c_uint64(episode), ctypes.c_uint32(purpose),
      ctypes.c_int(dim), _p(lo, ctypes.c_double), _p(hi, ctypes.c_double), _p(out, t))
    return out


def uniforms(seed, sample_id, episode, purpose, attempt, dim, dtype=np.float64):
    out = np.zeros(dim, dtype=dtype)
    if np.dtype(dtype) == np.float64:
        f, t = lib().oracle_uniforms_f64, ctypes.c_double
    else:
        f, t = lib().oracle_uniforms_f32, ctypes.c_float
    f(ctypes.c_uint64(seed), ctypes.c_uint64(sample_id), ctypes.c_uint64(episode), ctypes.c_uint32(purpose),
      ctypes.c_int(attempt), ctypes.c_int(dim), _p(out, t))
    return out


def random_actions(n, env_id0, seed, step):
    out = np.zeros(n, dtype=np.uint8)
    lib().oracle_random_actions(ctypes.c_int64(n), ctypes.c_uint64(env_id0), ctypes.c_uint64(seed),
                                ctypes.c_uint64(step), _p(out, ctypes.c_uint8))
    return out


def new_stats():
    return np.array([0.0, 0.0, 0.0, np.inf, -np.inf, 0.0])


def closed_loop(state, xi, elapsed, episode, seed, env_id0, tick0, K, max_steps=500, euler=True,
                actions=None, w=None, b=0.0, lo=None, hi=None, stats=None, log=False):
    """Run K steps of step -> TimeLimit -> auto-reset for every env (all arrays updated in place).

    Step k runs at clock ``tick0 + k`` (the framework's RNG contract: a reset is keyed by the tick it happens in).

    Returns dict(stats=..., done=(K,n) bool, truncated=(K,n) bool, states=(K,4,n)) (logs only if log=True).
    """
    assert state.dtype == np.float64 and xi.dtype == np.float64 and state.flags.c_contiguous and xi.flags.c_contiguous
    assert elapsed.dtype == np.int32 and episode.dtype == np.uint32
    n = state.shape[1]
    if stats is None:
        stats = new_stats()
    if actions is not None:
        actions = np.ascontiguousarray(actions, dtype=np.uint8)
        assert actions.shape == (K, n)
    w_ = _f64(w)
    lo_, hi_ = _f64(lo), _f64(hi)
    done = np.zeros((K, n), np.uint8) if log else None
    trunc = np.zeros((K, n), np.uint8) if log else None
    states = np.zeros((K, 4, n), np.float64) if log else None
    lib().oracle_closed_loop_f64(
        ctypes.c_int64(n), _p(state, ctypes.c_double), _p(xi, ctypes.c_double), _p(elapsed, ctypes.c_int32),
        _p(episode, ctypes.c_uint32), ctypes.c_uint64(seed), ctypes.c_uint64(env_id0), ctypes.c_uint64(tick0), ctypes.c_int(K),
        ctypes.c_int(max_steps), ctypes.c_int(1 if euler else 0), _p(actions, ctypes.c_uint8),
        _p(w_, ctypes.c_double), ctypes.c_double(b), _p(lo_, ctypes.c_double), _p(hi_, ctypes.c_double),
        _p(stats, ctypes.c_double), _p(done, ctypes.c_uint8), _p(trunc, ctypes.c_uint8), _p(states, ctypes.c_double))
    out = dict(stats=stats)
    if log:
        out.update(done=done.astype(bool), truncated=trunc.astype(bool), states=states)
    return out
