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


def _stats_worker(rank, world, port, out):
    """host logic of the multi-GPU update (uhc_b200/agent.py update_params): fp64 statistics riding an fp32 all-reduce exactly (base-2^18 digit planes),
    the ZFilter increments of every rank merged into one common running_state, global advantage moments."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from uhc_b200 import nn
    D = 7
    rng = np.random.RandomState(100 + rank)
    common = np.random.RandomState(5).normal(2.0, 3.0, (50, D))              # what every rank had agreed on before
    local = rng.normal(-1.0, 0.5, (30 + 10 * rank, D)) * (1 + rank)           # this rank's new observations

    def stats_of(x):
        return torch.tensor(np.concatenate([[len(x)], x.mean(0), ((x - x.mean(0)) ** 2).sum(0)]))
    z_sync = nn.zfilter_to_sums(stats_of(common), D)
    z_now = stats_of(np.concatenate([common, local]))                         # the rank-local running_state after its rollout
    adv = torch.tensor(rng.normal(0.3 * rank, 1.0 + rank, 1000))
    d = torch.cat([torch.stack([adv.sum(), (adv * adv).sum()]), torch.tensor([float(len(adv))]), nn.zfilter_to_sums(z_now, D) - z_sync])
    planes = nn.split_double(d)
    tail = torch.zeros(128, dtype=torch.float32)
    tail[:planes.numel()] = planes.reshape(-1)
    comm = nn.GradComm(world)
    comm.start(tail)                                                          # gloo: synchronous all-reduce(sum) of the fp32 buffer
    g = nn.join_double(tail[:planes.numel()].reshape(nn.SPLIT_CHUNKS, len(d)))
    merged = nn.zfilter_from_sums(z_sync + g[3:], D)
    gathered = [None] * world
    dist.all_gather_object(gathered, (local.tolist(), adv.tolist()))
    allx = np.concatenate([common] + [np.array(l) for l, _ in gathered])
    alla = np.concatenate([np.array(a) for _, a in gathered])
    ref = stats_of(allx)
    ok_z = bool(torch.allclose(merged, ref, rtol=1e-10, atol=1e-9))
    N, mean = g[2].item(), (g[0] / g[2]).item()
    var = ((g[1] - g[2] * mean * mean) / (g[2] - 1)).item()
    ok_a = abs(mean - alla.mean()) < 1e-9 and abs(var - alla.var(ddof=1)) < 1e-8 and N == len(alla)
    out.put((rank, ok_z, ok_a, merged.tolist(), comm.bytes))
    dist.destroy_process_group()


def test_statistics_tail_and_zfilter_merge_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_stats_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(30) for p in ps]
    assert all(r[1] and r[2] for r in res), res
    assert res[0][3] == res[1][3]                      # both ranks end with the SAME running_state
    assert res[0][4] == 128 * 4
