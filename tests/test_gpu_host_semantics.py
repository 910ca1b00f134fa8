"""Host-side semantics the round-1 review found missing: DR stream keyed by the constructor seed and the GLOBAL env id
(shard invariance of set_random_task), re-seeding restarts the episode stream, checkpoints carry the DR state, action
tensors are checked for device / shape, device-side errors surface at every synchronising call."""
import ctypes

import numpy as np
import pytest
import torch

import random_envs_b200 as renv

pytestmark = pytest.mark.gpu
SEARCH = [2.0, 20.0, 0.5, 3.0, 0.05, 0.3, 0.1, 1.0]


def _env(n, **kw):
    e = renv.RandomCartPoleVecEnv(n, **kw)
    e.set_dr_distribution("uniform", SEARCH)
    e.set_dr_training(True)
    return e


def test_set_random_task_is_independent_of_sharding_and_depends_on_the_seed():
    """Global env i gets the same xi whether it lives in one env of 1000 or in shard 1 of 2 (random_env.py:37-39 per
    env; ADVICE r1: the vector env used sample id 0 and DR seed 0 whatever the constructor said)."""
    whole = _env(1000, seed=7)
    lo, hi = _env(400, seed=7, env_id0=0), _env(600, seed=7, env_id0=400)
    for e in (whole, lo, hi):
        e.set_random_task()
    got = torch.cat([lo.get_task(), hi.get_task()])
    assert torch.equal(whole.get_task(), got)
    whole.set_random_task(); lo.set_random_task(); hi.set_random_task()          # second call: next Philox tick
    assert torch.equal(whole.get_task(), torch.cat([lo.get_task(), hi.get_task()])) and not torch.equal(whole.get_task(), got)
    other = _env(1000, seed=8); other.set_random_task(); other.set_random_task()
    assert not torch.equal(other.get_task(), whole.get_task())
    # the same through make_sharded_env's bookkeeping
    a = renv.make_sharded_env(1000, rank=1, world_size=2, seed=7)
    a.set_dr_distribution("uniform", SEARCH); a.set_random_task()
    b = _env(1000, seed=7); b.set_random_task()
    assert torch.equal(a.get_task(), b.get_task()[500:])


def test_reseeding_restarts_the_episode_stream():
    """env.seed(s); env.reset() is reproducible on the same object (the reference rebuilds np_random in seed())."""
    e = _env(2048, seed=1)
    e.seed(5); first = e.reset().clone(); xi = e.get_task().clone()
    act = e.sample_actions().clone()
    for _ in range(30):
        e.step(act)
    e.seed(5); again = e.reset().clone()
    assert torch.equal(first, again) and torch.equal(xi, e.get_task())
    g = renv.gym.make("RandomCartPole-v0")
    g.seed(3); s1 = g.reset(); g.step(0); g.step(1)
    g.seed(3); s2 = g.reset()
    assert np.array_equal(s1, s2)
    g.close()


def test_checkpoint_carries_the_dr_sampler_and_the_distribution():
    a = _env(3000, seed=11, max_episode_steps=20)
    a.reset()
    act = a.sample_actions().clone()
    for _ in range(25):
        a.step(act)
    a.set_random_task()                              # advances the sample_tasks call index
    sd = a.state_dict()
    b = renv.RandomCartPoleVecEnv(3000, seed=0, max_episode_steps=20)     # nothing configured by hand
    b.load_state_dict(sd)
    assert b.sampling == "uniform" and b.dr_training is True and np.array_equal(b.max_task, a.max_task)
    for _ in range(30):                              # auto-reset keeps resampling with the restored distribution
        a.step(act); b.step(act)
    assert torch.equal(a.obs, b.obs) and torch.equal(a.get_task(), b.get_task())
    a.set_random_task(); b.set_random_task()
    assert torch.equal(a.get_task(), b.get_task())   # same DR stream position
    assert np.array_equal(a.sample_tasks(5), b.sample_tasks(5))


def test_action_tensors_are_checked():
    e = _env(64, seed=2); e.reset()
    with pytest.raises(ValueError):
        e.step(torch.zeros(65, dtype=torch.uint8, device="cuda"))                   # wrong length, zero-copy candidate
    with pytest.raises(ValueError):
        e.step(torch.zeros((64, 1), dtype=torch.uint8, device="cuda"))
    e.step(torch.zeros(64, dtype=torch.uint8))                                       # CPU tensor: staged through a copy
    e.step(torch.ones(64, dtype=torch.int64, device="cuda"))
    e.check_dr_violations()
    e.step(torch.full((64,), 300, dtype=torch.int64, device="cuda"))                # would wrap to 44 in uint8: still invalid
    with pytest.raises(AssertionError, match="invalid"):
        e.allgather_stats()


def test_gaussian_failure_surfaces_at_every_synchronising_call():
    def failing():
        e = renv.RandomCartPoleVecEnv(512, seed=3, max_episode_steps=3)
        e.set_dr_distribution("gaussian", [-5.0, 0.1, 1.0, 0.1, 0.1, 0.01, 0.5, 0.05]); e.set_dr_training(True)
        e.reset()
        return e
    for call in (lambda e: e.episode_stats(), lambda e: e.state_dict(), lambda e: e.allgather_stats(),
                 lambda e: e.step_host(np.zeros(512, dtype=np.uint8)), lambda e: e.check_dr_violations()):
        e = failing()
        with pytest.raises(Exception, match="Not all samples were above > 0.1 after 2 attempts"):
            call(e)
        e.set_dr_training(False)
        e.check_dr_violations()                      # consumed: the next check is clean
    # an earlier failure is not blamed on an unrelated sample_tasks call, and is still reported afterwards
    e = failing()
    ok = renv.TaskSampler("RandomCartPole-v0")
    e.set_dr_distribution("uniform", SEARCH)
    assert e.sample_tasks(4).shape == (4, 4)
    with pytest.raises(Exception, match="Not all samples were above"):
        e.check_dr_violations()
    del ok


@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, 4096, 100003])
def test_pack_flags_matches_numpy_packbits(n):
    """renv_pack_flags_u8: bit (i & 7) of byte i >> 3 == flags[i] != 0 (numpy bitorder "little"), any n, flags past n
    ignored, the 16-byte-aligned vector path and the unaligned scalar path."""
    from random_envs_b200 import _device, _lib
    rng = np.random.default_rng(n)
    flags = rng.integers(0, 3, size=n + 4, dtype=np.uint8)          # 0, 1, 2: any non-zero byte is a set flag
    flags[rng.integers(0, n)] = 255
    want = np.packbits(flags[:n] != 0, bitorder="little")
    dev = torch.as_tensor(flags, device="cuda")
    words = (n + 31) // 32
    for offset in (0, 4):                                           # offset 4: aligned to 4 only -> scalar path
        src = dev[offset:offset + n]
        want_o = np.packbits(flags[offset:offset + n] != 0, bitorder="little")
        bits = torch.full((words,), -1, dtype=torch.int32, device="cuda")
        _lib.call("renv_pack_flags_u8", _device.ptr(src), _device.ptr(bits), n, _device.stream_ptr(dev.device))
        got = bits.cpu().numpy().view(np.uint8)[:len(want_o)]
        assert np.array_equal(got, want_o)
        assert np.array_equal(np.unpackbits(got, count=n, bitorder="little"), (flags[offset:offset + n] != 0).astype(np.uint8))
    assert np.array_equal(np.packbits(flags[:n] != 0, bitorder="little"), want)
    with pytest.raises(_lib.RenvError):
        _lib.call("renv_pack_flags_u8", ctypes.c_void_p(dev.data_ptr() + 1), _device.ptr(bits), n,
                  _device.stream_ptr(dev.device))
    with pytest.raises(_lib.RenvError):
        _lib.call("renv_pack_flags_u8", None, _device.ptr(bits), n, _device.stream_ptr(dev.device))


@pytest.mark.parametrize("n", [1000, 4097])
def test_step_host_with_packed_flags_equals_byte_flags(n):
    """step_host_* moves done / truncated as bits by default; the numpy results equal those of the byte path on every
    step, including the steps where episodes end (TimeLimit 20 so that `truncated` fires too)."""
    a = _env(n, seed=11, max_episode_steps=20)
    b = _env(n, seed=11, max_episode_steps=20, pack_host_flags=False)
    assert a.host_bytes_per_step()[1] < b.host_bytes_per_step()[1]
    a.reset(); b.reset()
    rng = np.random.default_rng(0)
    ended = truncated = 0
    for _ in range(60):
        act = rng.integers(0, 2, size=n, dtype=np.uint8)
        ra, rb = a.step_host(act), b.step_host(act)
        for x, y in zip(ra, rb):
            assert np.array_equal(x, y)
        ended += int(ra[2].sum()); truncated += int(ra[3].sum())
    assert ended > n and truncated > 0
