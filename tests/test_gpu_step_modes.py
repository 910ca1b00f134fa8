"""Step-kernel modes against the plain grid-ordered 62-byte step (which the rest of the suite pins to the oracle): the
tile-granular step ordering (overlapping launches, also under CUDA-graph replay), the lean 54-byte step, and the two
error conditions detected on the device (invalid action, gaussian DR failure)."""

import numpy as np
import pytest
import torch

import random_envs_b200 as renv

pytestmark = pytest.mark.gpu
SEARCH = [2.0, 20.0, 0.5, 3.0, 0.05, 0.3, 0.1, 1.0]


def _env(n, lean=False, **kw):
    env = renv.RandomCartPoleVecEnv(n, dtype="float32", seed=11, env_id0=12345, max_episode_steps=25, lean=lean, **kw)
    env.set_dr_distribution("uniform", SEARCH)
    env.set_dr_training(True)
    env.reset()
    return env


def _snapshot(env, out):
    obs, reward, done, info = out
    return [obs.clone(), reward.clone(), done.clone(), info["TimeLimit.truncated"].clone(), env.get_task().clone(),
            env.elapsed.clone(), env.episode.clone()]


@pytest.mark.parametrize("n,dtype", [(3 * 1024 + 5, "float32"), (1 << 20, "float32"), (300000, "float64")])
def test_tile_ordered_steps_equal_grid_ordered_steps(n, dtype):
    """Back-to-back launches WITHOUT any host sync in between (the overlapping case), several env batches round-robin on
    one stream, then the same K steps as a replayed CUDA graph: every buffer identical to the grid-ordered run."""
    def build(tile):
        envs = []
        for b in range(3):
            e = renv.RandomCartPoleVecEnv(n, dtype=dtype, seed=11, env_id0=b * n, max_episode_steps=25, tile_ordering=tile)
            e.set_dr_distribution("uniform", SEARCH); e.set_dr_training(True); e.reset()
            envs.append(e)
        return envs
    ref, til = build(False), build(True)
    acts = [ref[0].sample_actions().clone() for _ in range(4)]
    K = 45
    for k in range(K):
        for envs in (ref, til):
            envs[k % 3].step(acts[k % 4])
    torch.cuda.synchronize()
    for a, b in zip(ref, til):
        assert torch.equal(a.state, b.state) and torch.equal(a.get_task(), b.get_task())
        assert torch.equal(a.elapsed, b.elapsed) and torch.equal(a.episode, b.episode)
        b.check_dr_violations()
    # CUDA-graph replay (tickets are taken on the device: nothing about the launch order is baked into the graph;
    # the Philox tick IS a launch parameter, so the eager twin replays the same ticks)
    tick0 = [e._tick for e in til]
    assert tick0 == [e._tick for e in ref]
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for k in range(12):
                til[k % 3].step(acts[k % 4])
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(3):
        g.replay()
    for _ in range(3):
        for e, t0 in zip(ref, tick0):
            e._tick = t0
        for k in range(12):
            ref[k % 3].step(acts[k % 4])
    torch.cuda.synchronize()
    for a, b in zip(ref, til):
        assert torch.equal(a.state, b.state) and torch.equal(a.get_task(), b.get_task()) and torch.equal(a.elapsed, b.elapsed)
        b.check_dr_violations()


def test_tile_ordering_survives_interleaved_resets_rollouts_and_a_corrupted_progress_array():
    n = 5000
    a = renv.RandomCartPoleVecEnv(n, dtype="float32", seed=5, max_episode_steps=30, tile_ordering=False)
    b = renv.RandomCartPoleVecEnv(n, dtype="float32", seed=5, max_episode_steps=30, tile_ordering=True)
    for e in (a, b):
        e.set_dr_distribution("uniform", SEARCH); e.set_dr_training(True); e.reset()
    act = a.sample_actions().clone()
    mask = (torch.arange(n, device="cuda") % 7 == 0).to(torch.uint8)
    for e in (a, b):
        for _ in range(5):
            e.step(act)
        e.reset(mask)
        e.step(act)
        e.rollout((0.0, 0.0, 1.0, 0.0), 0.0, 17)
        for _ in range(5):
            e.step(act)
    assert torch.equal(a.state, b.state) and torch.equal(a.elapsed, b.elapsed)
    # a progress array somebody scribbled on: the step still runs (bounded wait), the host is told, the protocol heals
    b._buffers["progress"][1] = 12345
    b.step(act)
    with pytest.raises(RuntimeError, match="timed out"):
        b.episode_stats()
    a.step(act)
    b.step(act); a.step(act)
    b.check_dr_violations()
    assert torch.equal(a.state, b.state)


@pytest.mark.parametrize("n", [4099, 148 * 512 + 300, 1 << 20])
def test_lean_step_matches_the_62_byte_step(n):
    a, b = _env(n), _env(n, lean=True)
    for k in range(40):
        act = a.sample_actions().clone()
        oa, ra, da, ia = a.step(act)
        ob, rb, db, ib = b.step(act)
        assert torch.equal(oa, ob) and torch.equal(da, db), "step %d" % k
        assert torch.equal(ia["TimeLimit.truncated"], ib["TimeLimit.truncated"])
        assert torch.equal(a.elapsed, b.elapsed) and torch.equal(a.get_task(), b.get_task())
        assert torch.equal(ra, rb) and float(rb.min()) == 1.0 == float(rb.max())     # the reward is the constant 1.0
    with pytest.raises(ValueError):
        b.rollout((0, 0, 1, 0), 0.0, 5)


def test_lean_needs_fp32_auto_reset():
    for kw in (dict(dtype="float64"), dict(auto_reset=False), dict(noisy=True), dict(max_episode_steps=70000)):
        with pytest.raises(ValueError):
            renv.RandomCartPoleVecEnv(64, lean=True, **kw)


@pytest.mark.parametrize("n", [1000, 148 * 512])
def test_invalid_action_is_flagged_on_the_device_and_raised_at_the_next_sync(n):
    """random_cartpole.py:173-174 asserts inside step(); the vector env raises the same AssertionError at its next
    synchronising call (step itself never blocks)."""
    env = _env(n)
    act = torch.zeros(n, dtype=torch.uint8, device="cuda")
    env.step(act)
    env.check_dr_violations()                  # clean
    act[n // 2] = 2
    env.step(act)
    with pytest.raises(AssertionError, match="invalid"):
        env.episode_stats()
    env.check_dr_violations()                  # the flag was consumed
    # host-buffer path: raised by step_host_wait of the same step; wide integer inputs keep their invalidity
    bad = np.zeros(n, dtype=np.int64); bad[3] = 256
    with pytest.raises(AssertionError, match="invalid"):
        env.step_host(bad)
    quiet = _env(n, validate_actions=False)
    quiet.step(act)
    quiet.episode_stats()                      # validate_actions=False: pushed left, nothing raised
    with pytest.raises(AssertionError):
        env.step(torch.zeros(n, dtype=torch.float32, device="cuda"))


def test_gaussian_failure_is_raised_on_the_vector_path():
    """random_env.py:181-186 via set_random_task on reset: mean far below 0.1 -> every draw fails three times."""
    env = renv.RandomCartPoleVecEnv(2048, dtype="float32", seed=3, max_episode_steps=5)
    env.set_dr_distribution("gaussian", [-5.0, 0.1, 1.0, 0.1, 0.1, 0.01, 0.5, 0.05])
    env.set_dr_training(True)
    env.reset()
    with pytest.raises(Exception, match="Not all samples were above > 0.1 after 2 attempts"):
        env.state_dict()
    env2 = renv.RandomCartPoleVecEnv(2048, dtype="float32", seed=3, max_episode_steps=5)
    env2.set_dr_distribution("gaussian", [-5.0, 0.1, 1.0, 0.1, 0.1, 0.01, 0.5, 0.05])
    with pytest.raises(Exception, match="Not all samples were above"):
        env2.set_random_task()
    env2.set_dr_training(True)
    env2.reset()
    env2.check_dr_violations.__func__          # (exists)
    with pytest.raises(Exception, match="Not all samples were above"):
        env2.step_host(np.zeros(2048, dtype=np.uint8))
