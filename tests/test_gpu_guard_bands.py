"""Out-of-bounds WRITE check without compute-sanitizer (closed on this pool): every buffer handed to the C ABI is a
slice of one arena with canary bands on both sides, sized with the tightest `ld` the contract allows; after reset /
step / noisy step / rollout / samplers at awkward sizes every band must still hold its pattern."""
import ctypes
import math

import numpy as np
import pytest
import torch

from random_envs_b200 import _device, _lib

pytestmark = pytest.mark.gpu

BAND = 256          # bytes of canary on each side of every buffer
PATTERN = 0xA5


class Arena:
    def __init__(self, nbytes):
        self.buf = torch.full((nbytes,), PATTERN, dtype=torch.uint8, device="cuda")
        self.off = 0
        self.live = []

    def take(self, nbytes, dtype):
        self.off = (self.off + 255) // 256 * 256 + BAND
        start = self.off
        self.off += nbytes
        self.live.append((start, nbytes))
        view = self.buf[start:start + nbytes].view(dtype)
        self.off += BAND
        return view

    def check(self):
        mask = torch.ones_like(self.buf, dtype=torch.bool)
        for start, nbytes in self.live:
            mask[start:start + nbytes] = False
        assert bool((self.buf[mask] == PATTERN).all()), "a kernel wrote outside its buffer"


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("n", [1, 2, 3, 5, 255, 257, 1031])
def test_cartpole_kernels_stay_inside_their_buffers(dtype, n):
    lib = _lib.load()
    suffix = "f32" if dtype == torch.float32 else "f64"
    esz = 4 if dtype == torch.float32 else 8
    V = 16 // esz
    ld = (n + V - 1) // V * V                       # tightest legal stride
    ar = Arena(1 << 20)
    state = ar.take(4 * ld * esz, dtype); obs = ar.take(4 * ld * esz, dtype)
    xi = ar.take(4 * ld * esz, dtype)
    elapsed = ar.take(ld * 4, torch.int32); episode = ar.take(ld * 4, torch.int32); beyond = ar.take(ld * 4, torch.int32)
    reward = ar.take(ld * esz, dtype)
    done = ar.take((n + 3) // 4 * 4, torch.uint8); trunc = ar.take((n + 3) // 4 * 4, torch.uint8)
    action = ar.take((n + 15) // 16 * 16, torch.uint8); mask = ar.take((n + 3) // 4 * 4, torch.uint8)
    stats = ar.take(6 * 8, torch.float64); viol = ar.take(8, torch.int64)
    for t in (state, obs, xi, reward):
        t.fill_(1.0)
    for t in (elapsed, episode, done, trunc, action, viol):
        t.zero_()
    beyond.fill_(-1); mask.fill_(1)
    stats.copy_(torch.tensor([0.0, 0.0, 0.0, math.inf, -math.inf, 0.0], dtype=torch.float64))
    env = _lib.CartpoleEnv()
    env.state, env.xi, env.elapsed, env.episode, env.beyond = (state.data_ptr(), xi.data_ptr(), elapsed.data_ptr(),
                                                               episode.data_ptr(), beyond.data_ptr())
    env.n, env.ld, env.env_id0, env.seed = n, ld, 1 << 33, 99
    noise = _lib.ObsNoise(); noise.obs, noise.std = obs.data_ptr(), 0.01
    stream = _device.stream_ptr(torch.device("cuda", 0))
    p = _device.ptr
    w = (ctypes.c_double * 4)(0.0, 0.0, 1.0, 0.0)
    for dr_type, a, b in (("uniform", [2.0, 0.5, 0.05, 0.1], [20.0, 3.0, 0.3, 1.0]),
                          ("truncnorm", [9.8, 1.0, 0.5, 0.5], [1.0, 0.5, 0.3, 0.2]),
                          ("gaussian", [9.8, 1.0, 0.3, 0.5], [1.0, 0.1, 0.05, 0.05])):
        cfg = _lib.make_dr_cfg(dr_type, a, b, [0.1] * 4)
        tick = 0
        _lib.call("renv_cartpole_reset_" + suffix, ctypes.byref(env), p(mask), tick, ctypes.byref(cfg), p(viol), stream)
        _lib.call("renv_cartpole_reset_noisy_" + suffix, ctypes.byref(env), ctypes.byref(noise), None, tick + 1,
                  ctypes.byref(cfg), p(viol), stream)
        tick = 2
        for auto_reset in (1, 0):
            for k in range(12):
                _lib.call("renv_random_actions_u8", p(action), n, env.env_id0, env.seed, tick, stream)
                _lib.call("renv_cartpole_step_" + suffix, ctypes.byref(env), p(action), p(reward), p(done), p(trunc),
                          _lib.EULER, 7, auto_reset, tick, ctypes.byref(cfg), p(viol), stream)
                _lib.call("renv_cartpole_step_noisy_" + suffix, ctypes.byref(env), ctypes.byref(noise), p(action), p(reward),
                          p(done), None, _lib.SEMI_IMPLICIT, 7, auto_reset, tick + 1, ctypes.byref(cfg), p(viol), stream)
                tick += 2
        _lib.call("renv_cartpole_reset_" + suffix, ctypes.byref(env), None, tick, ctypes.byref(cfg), p(viol), stream)
        for K in (1, 2, 37):
            _lib.call("renv_cartpole_rollout_" + suffix, ctypes.byref(env), w, 0.0, K, _lib.EULER, 11, tick + 1,
                      ctypes.byref(cfg), p(stats), p(viol), stream)
            tick += K
            _lib.call("renv_cartpole_rollout_noisy_" + suffix, ctypes.byref(env), ctypes.byref(noise), w, 0.0, K, _lib.SEMI_IMPLICIT,
                      11, tick + 1, ctypes.byref(cfg), p(stats), p(viol), stream)
            tick += K
    torch.cuda.synchronize()
    ar.check()
    assert bool(torch.isfinite(state.view(4, ld)[:, :n]).all()) and bool((elapsed[:n] >= 0).all())


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("dim", [1, 3, 4, 5, 8, 9, 13, 23, 30, 32])
def test_sampler_kernels_stay_inside_their_buffers(dtype, dim):
    esz = 4 if dtype == torch.float32 else 8
    stream = _device.stream_ptr(torch.device("cuda", 0))
    for n in (1, 7, 129, 2049, 4099):
        ar = Arena(4 << 20)
        out = ar.take(n * dim * esz, dtype); viol = ar.take(8, torch.int64)
        viol.zero_()
        for dr_type in ("uniform", "truncnorm", "gaussian", "fullgaussian"):
            if dr_type == "fullgaussian":
                cfg = _lib.make_dr_cfg(dr_type, [2.0] * dim, [0.5] * dim, [10.0] * dim, factor=np.eye(dim) * 0.5)
            else:
                cfg = _lib.make_dr_cfg(dr_type, [1.0] * dim, [2.0 if dr_type == "uniform" else 0.1] * dim, [0.1] * dim)
            out.fill_(-1.0)
            _lib.call("renv_dr_sample_f32" if esz == 4 else "renv_dr_sample_f64", _device.ptr(out), n, ctypes.byref(cfg), 5, 0,
                      0, _device.ptr(viol), stream)
            torch.cuda.synchronize()
            ar.check()
            assert bool((out > 0).all()), (dr_type, n, dim)       # every element was written


@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, 4097])
def test_pack_flags_stays_inside_its_buffers(n):
    """renv_pack_flags_u8 writes exactly ceil(n / 32) words and reads nothing it may not (flags sized n rounded up to 4
    bytes only): canaries around both buffers stay intact and flags past n never reach the bits."""
    ar = Arena(1 << 16)
    flags = ar.take((n + 3) // 4 * 4, torch.uint8)
    bits = ar.take((n + 31) // 32 * 4, torch.int32)
    flags.fill_(1)                                   # incl. the padding bytes past n
    bits.zero_()
    _lib.call("renv_pack_flags_u8", _device.ptr(flags), _device.ptr(bits), n, _device.stream_ptr(torch.device("cuda", 0)))
    torch.cuda.synchronize()
    ar.check()
    got = np.unpackbits(bits.cpu().numpy().view(np.uint8), bitorder="little")
    assert got[:n].all() and not got[n:].any()
