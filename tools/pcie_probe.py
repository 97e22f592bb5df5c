#!/usr/bin/env python
"""Pinned host<->device copy bandwidth of this box, per GPU and with ALL ranks copying at once (context for the e2e
number: the end-to-end path moves every image down and every result up through host memory).

    python tools/pcie_probe.py                                     # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe.py

Prints one JSON line (rank 0): per-rank H2D / D2H / simultaneous GB/s measured while every rank copies (wall clock
between barriers), and their sums.
"""
import json
import os
import time

import torch

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
flag = torch.zeros(1, device="cuda")


def barrier():
    if world > 1:
        dist.all_reduce(flag)
    torch.cuda.synchronize()


def timed(fn, reps=6):
    for _ in range(2):
        fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    return reps * n / dt / 1e9


def up():
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)


def down():
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


def both():
    up(); down()


res = dict(h2d=timed(up), d2h=timed(down), bidir_each=timed(both))
vals = torch.tensor([res["h2d"], res["d2h"], res["bidir_each"]], device="cuda")
if world > 1:
    allv = [torch.zeros_like(vals) for _ in range(world)]
    dist.all_gather(allv, vals)
    allv = torch.stack(allv).cpu().tolist()
else:
    allv = [vals.cpu().tolist()]
if rank == 0:
    out = dict(ranks=world, host_cores=os.cpu_count(),
               per_rank_gbs=[dict(h2d=round(a, 2), d2h=round(b, 2), bidir_each_direction=round(c, 2)) for a, b, c in allv],
               aggregate_gbs=dict(h2d=round(sum(v[0] for v in allv), 1), d2h=round(sum(v[1] for v in allv), 1),
                                  bidir_total=round(2 * sum(v[2] for v in allv), 1)))
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
