"""T5: two ranks, one per GPU, NCCL.  Skipped unless >= 2 CUDA devices are visible (gpurun --gpus 2).

Per-env trajectories of a sharded job must equal the single-GPU run (Philox is keyed by the GLOBAL env id) and the
all-gathered return statistics must equal the single-GPU reduction.
"""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SEARCH = [2.0, 20.0, 0.5, 3.0, 0.05, 0.3, 0.1, 1.0]
N_TOTAL, K = 1 << 16, 200
W = (0.0, 0.0, 1.0, 0.0)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    import random_envs_b200 as renv
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        env = renv.make_sharded_env(N_TOTAL, dtype="float32", seed=21, device="cuda:%d" % rank)
        env.set_dr_distribution("uniform", SEARCH); env.set_dr_training(True)
        env.reset()
        env.rollout(W, 0.0, K)
        combined, gathered = renv.allgather_stats(env.stats_tensor)
        np.save(os.path.join(out_dir, "obs_%d.npy" % rank), env.obs.cpu().numpy())
        np.save(os.path.join(out_dir, "xi_%d.npy" % rank), env.get_task().cpu().numpy())
        np.save(os.path.join(out_dir, "stats_%d.npy" % rank), combined.cpu().numpy())
        np.save(os.path.join(out_dir, "range_%d.npy" % rank), np.array([env.env_id0, env.num_envs]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_shards_match_single_gpu(tmp_path):
    import torch.multiprocessing as mp
    import random_envs_b200 as renv
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    whole = renv.RandomCartPoleVecEnv(N_TOTAL, dtype="float32", seed=21, device="cuda:0")
    whole.set_dr_distribution("uniform", SEARCH); whole.set_dr_training(True)
    whole.reset(); whole.rollout(W, 0.0, K)
    obs = np.concatenate([np.load(tmp_path / ("obs_%d.npy" % r)) for r in range(2)])
    xi = np.concatenate([np.load(tmp_path / ("xi_%d.npy" % r)) for r in range(2)])
    assert np.array_equal(obs, whole.obs.cpu().numpy()) and np.array_equal(xi, whole.get_task().cpu().numpy())
    s0, s1 = np.load(tmp_path / "stats_0.npy"), np.load(tmp_path / "stats_1.npy")
    assert np.array_equal(s0, s1) and np.array_equal(s0, whole.stats_tensor.cpu().numpy())
    r0, r1 = np.load(tmp_path / "range_0.npy"), np.load(tmp_path / "range_1.npy")
    assert list(r0) == [0, N_TOTAL // 2] and list(r1) == [N_TOTAL // 2, N_TOTAL // 2]
