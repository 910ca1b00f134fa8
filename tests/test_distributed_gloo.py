"""N>1 host logic on CPU: world_size-2 gloo processes exercise sharding + the stats all-gather."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from random_envs_b200 import allgather_stats, combine_stats, shard_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        start, stop = shard_range(total, rank, world)
        # deterministic fake per-rank statistics derived from the owned env ids
        ids = np.arange(start, stop, dtype=np.float64)
        returns = 1.0 + (ids % 500)
        local = torch.tensor([len(ids), returns.sum(), (returns ** 2).sum(), returns.min(), returns.max(), returns.sum()],
                             dtype=torch.float64)
        combined, gathered = allgather_stats(local)
        assert gathered.shape == (world, 6)
        assert torch.equal(gathered[rank], local)
        np.save(os.path.join(out_dir, "combined_%d.npy" % rank), combined.numpy())
        np.save(os.path.join(out_dir, "gathered_%d.npy" % rank), gathered.numpy())
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_allgather_matches_single_process(tmp_path):
    total, world = 100003, 2
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    ids = np.arange(total, dtype=np.float64)
    returns = 1.0 + (ids % 500)
    want = np.array([total, returns.sum(), (returns ** 2).sum(), returns.min(), returns.max(), returns.sum()])
    c0 = np.load(tmp_path / "combined_0.npy"); c1 = np.load(tmp_path / "combined_1.npy")
    assert np.array_equal(c0, c1)                      # every rank ends with the same combined vector
    assert np.allclose(c0, want, rtol=1e-14, atol=0)
    g0 = np.load(tmp_path / "gathered_0.npy")
    assert np.allclose(combine_stats(g0), want, rtol=1e-14, atol=0)
    assert g0[0, 0] + g0[1, 0] == total


def test_allgather_is_identity_without_process_group():
    local = torch.tensor([3.0, 30.0, 400.0, 5.0, 15.0, 30.0], dtype=torch.float64)
    combined, gathered = allgather_stats(local)
    assert torch.equal(combined, local) and gathered.shape == (1, 6)
