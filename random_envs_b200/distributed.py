"""Multi-GPU layer: contiguous env-id sharding and the one collective on the path.

Envs are independent, so the step path has NO collective: rank r of G owns the
global env ids [r*N/G, (r+1)*N/G) and, because every Philox draw is keyed by the GLOBAL id, the
trajectories of env i are identical whatever G is.  The only exchange is one ``all_gather`` of the
6-double return-statistics vector per rollout iteration (NCCL over NVLink/NVSwitch on GPUs, gloo in
the CPU tests), combined on every rank with (sum, sum, sum, min, max, sum).
"""
import math

import numpy as np

STAT_NAMES = ("episodes", "sum_return", "sum_return_sq", "min_return", "max_return", "sum_length")


def shard_range(num_envs_total, rank, world_size):
    """Contiguous [start, stop) of global env ids owned by ``rank``; remainders go to the low ranks."""
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, rem = divmod(int(num_envs_total), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def combine_stats(per_rank):
    """Reduce an (R, 6) array of per-rank statistics to one 6-vector."""
    a = np.asarray(per_rank, dtype=np.float64).reshape(-1, len(STAT_NAMES))
    return np.array([a[:, 0].sum(), a[:, 1].sum(), a[:, 2].sum(), a[:, 3].min(), a[:, 4].max(), a[:, 5].sum()])


def summarize_stats(stats):
    s = np.asarray(stats, dtype=np.float64)
    n = s[0]
    mean = s[1] / n if n > 0 else math.nan
    var = max(s[2] / n - mean * mean, 0.0) if n > 0 else math.nan
    return dict(episodes=int(n), mean_return=mean, std_return=math.sqrt(var) if n > 0 else math.nan,
                min_return=s[3] if n > 0 else math.nan, max_return=s[4] if n > 0 else math.nan,
                mean_length=s[5] / n if n > 0 else math.nan)


def allgather_stats(stats_tensor, group=None):
    """All-gather the per-rank stats tensor (6 float64, on the backend's device) and combine.

    Returns (combined 6-vector tensor on the same device, gathered (world, 6) tensor).  With
    ``torch.distributed`` uninitialised this is the identity (single process).
    """
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return stats_tensor.clone(), stats_tensor.reshape(1, -1).clone()
    world = dist.get_world_size(group)
    flat = torch.empty(world * stats_tensor.numel(), dtype=stats_tensor.dtype, device=stats_tensor.device)
    dist.all_gather_into_tensor(flat, stats_tensor.contiguous().reshape(-1), group=group)
    gathered = flat.reshape(world, stats_tensor.numel())
    combined = torch.stack([gathered[:, 0].sum(), gathered[:, 1].sum(), gathered[:, 2].sum(), gathered[:, 3].min(),
                            gathered[:, 4].max(), gathered[:, 5].sum()])
    return combined, gathered


def make_sharded_env(num_envs_total, rank=None, world_size=None, **kwargs):
    """RandomCartPoleVecEnv holding this rank's shard; rank/world default to torch.distributed's."""
    from .vector_env import RandomCartPoleVecEnv
    if rank is None or world_size is None:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            rank, world_size = dist.get_rank(), dist.get_world_size()
        else:
            rank, world_size = 0, 1
    start, stop = shard_range(num_envs_total, rank, world_size)
    return RandomCartPoleVecEnv(stop - start, env_id0=start, **kwargs)
