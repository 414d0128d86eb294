"""CPU, world_size 2, gloo: the N>1 host logic -- gradient all-reduce wrapper and per-rank clip sampling streams."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from uhc_b200.agent import ClipSampler, make_nccl_grad_sync
    sync = make_nccl_grad_sync(world)
    g = [torch.full((3, 4), float(rank + 1)), torch.arange(5, dtype=torch.float32) * (rank + 1)]
    r = sync(g)
    ok = torch.allclose(r[0], torch.full((3, 4), 1.5)) and torch.allclose(r[1], torch.arange(5, dtype=torch.float32) * 1.5)
    s = ClipSampler(np.array([100, 200, 300]), t_min=5, t_max=60, seed=1 * 9973 + rank)
    clip, start, length = s.sample(64)
    valid = bool(((start >= 0) & (start < np.array([100, 200, 300])[clip] - 5) & (length <= 60) & (length >= 5)).all())
    gathered = [None] * world
    dist.all_gather_object(gathered, start.tolist())
    out.put((rank, bool(ok), valid, gathered[0] != gathered[1]))
    dist.destroy_process_group()


def test_grad_sync_and_sampler_streams_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in ps]
    [p.join(30) for p in ps]
    assert all(r[1] and r[2] and r[3] for r in res), res
