"""Image-sharded execution of the pseudo-label path over a pool.

Every image is independent (the loop body of functions.py:2844 carries no cross-image
state except the ``im_sizes`` dict), so the pool shards by image index ``i % world`` with
one process per GPU and NO collective on the data path.  The single exchange is the
coverage statistic of row a10: an all-reduce of ``int64[3] = {sum im_size, sum pred_size,
count}`` (NCCL over NVLink on GPUs, gloo in the CPU tests), after which every rank can
form ``mean_im_size = round(sum / count, 0)`` exactly as functions.py:2889 does.
"""
from __future__ import annotations

import os

import numpy as np

__all__ = ["shard_indices", "shard_bounds", "allreduce_stats", "mean_im_size", "dist_env"]


def dist_env():
    """(rank, local_rank, world) from the torchrun environment; (0, 0, 1) when absent."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)))


def shard_indices(n_items, rank, world):
    """Indices of the pool owned by ``rank``: ``i % world == rank``."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    return np.arange(rank, n_items, world, dtype=np.int64)


def shard_bounds(n_items, rank, world):
    """Size of every rank's shard under the modulo partition (they differ by at most one)."""
    base, extra = divmod(n_items, world)
    return base + (1 if rank < extra else 0)


def allreduce_stats(sum_im_size, sum_pred_size, count, device=None):
    """Sum ``{sum im_size, sum pred_size, count}`` over all ranks (int64: order-independent, bit-reproducible).

    Works without an initialised process group (single process) and with any backend."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(sum_im_size), int(sum_pred_size), int(count)], dtype=torch.int64,
                     device=device if device is not None else "cpu")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    s = t.cpu().tolist()
    return int(s[0]), int(s[1]), int(s[2])


def mean_im_size(sum_im_size, count):
    """functions.py:2889: Python ``round(sum / len, 0)`` (banker's rounding) -> float."""
    return round(int(sum_im_size) / int(count), 0)
