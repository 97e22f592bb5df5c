"""N > 1 host logic on CPU: image sharding by i % world and the int64[3] statistics
all-reduce that reproduces mean_im_size (functions.py:2889), world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from inconsistencymasks_b200 import pool
from oracle import ref_im


def test_shards_partition_the_pool():
    for n in (0, 1, 7, 8, 1000, 1001):
        for world in (1, 2, 4, 8):
            parts = [pool.shard_indices(n, r, world) for r in range(world)]
            allidx = np.sort(np.concatenate(parts)) if n else np.array([], np.int64)
            assert np.array_equal(allidx, np.arange(n))
            for r in range(world):
                assert len(parts[r]) == pool.shard_bounds(n, r, world)
    with pytest.raises(ValueError):
        pool.shard_indices(10, 2, 2)


def test_single_process_stats():
    assert pool.allreduce_stats(5, 7, 2) == (5, 7, 2)
    assert pool.mean_im_size(5, 2) == 2.0      # 2.5 -> 2 (banker's)
    assert pool.mean_im_size(3, 2) == 2.0      # 1.5 -> 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, sizes, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = pool.shard_indices(len(sizes), rank, world)
        local = int(np.asarray(sizes)[mine].sum())
        tot, _, cnt = pool.allreduce_stats(local, 0, len(mine))
        out[rank] = (tot, cnt, pool.mean_im_size(tot, cnt))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_mean_im_size():
    rng = np.random.default_rng(3)
    sizes = rng.integers(0, 5000, size=101).tolist()
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), sizes, out), nprocs=world, join=True)
    expect = ref_im.mean_im_size(sizes)
    for r in range(world):
        tot, cnt, mean = out[r]
        assert tot == sum(sizes) and cnt == len(sizes)
        assert mean == expect
