"""N>1 host logic on CPU: world_size-2 gloo job — frame sharding is a partition, timings reduce with MAX."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fullysparsefusion_b200 import dist as fdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = fdist.frame_ids(rank, world, 7)
    t = torch.tensor([10.0 + rank, 5.0 - rank], dtype=torch.float64)
    fdist.max_over_ranks(t)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    # training-side collectives: one bucket for k reductions, reduce_mean for several scalars, naive sync-BN statistics
    a, b = torch.full((3,), float(rank + 1)), torch.arange(4, dtype=torch.float32).view(2, 2) * (rank + 1)
    fdist.coalesced_all_reduce([a, b])
    scal = [torch.tensor(2.0 * rank), torch.tensor(10.0 + rank, dtype=torch.float64)]
    means = fdist.reduce_means(scal)
    x = torch.arange(6, dtype=torch.float32).view(3, 2) + 10 * rank if rank == 0 else torch.ones(5, 2) * 4
    mu, var = fdist.naive_sync_bn_stats(x)
    extra = dict(a=a.tolist(), b=b.tolist(), means=[float(m) for m in means], scal=[float(v) for v in scal],
                 dtypes=[str(m.dtype) for m in means], mu=mu.tolist(), var=var.tolist())
    if rank == 0:
        out.put((gathered, t.tolist(), [fdist.frame_seed(r, i) for r in range(world) for i in range(3)], extra))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_max_timing():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, t, seeds, extra = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(sum(gathered, [])) == list(range(7))            # a partition of the frames
    assert gathered[0] == [0, 2, 4, 6] and gathered[1] == [1, 3, 5]
    assert t == [11.0, 5.0]                                        # MAX over ranks
    assert len(set(seeds)) == len(seeds)                           # disjoint synthetic streams
    assert fdist.throughput(10, 2, 0.5) == 40.0
    assert extra["a"] == [3.0] * 3 and extra["b"] == [[0.0, 3.0], [6.0, 9.0]]        # sums over the two ranks, shapes kept
    assert extra["means"] == [1.0, 10.5] and extra["scal"] == [0.0, 10.0]            # means over ranks; inputs untouched
    assert extra["dtypes"] == ["torch.float32", "torch.float64"]
    # rank 0 rows [[0,1],[2,3],[4,5]] (mean [2,3], mean sq [20/3, 35/3]); rank 1 all 4 (mean 4, mean sq 16): equal rank weights
    import numpy as np
    np.testing.assert_allclose(extra["mu"], [3.0, 3.5], rtol=1e-6)
    np.testing.assert_allclose(extra["var"], [(20 / 3 + 16) / 2 - 9.0, (35 / 3 + 16) / 2 - 12.25], rtol=1e-5)


def test_single_process_is_identity():
    t = torch.tensor([3.0])
    assert fdist.max_over_ranks(t).item() == 3.0
    assert fdist.frame_ids(0, 1, 3) == [0, 1, 2]
    v = [torch.tensor(2.0), torch.tensor(5.0)]
    assert [float(m) for m in fdist.reduce_means(v)] == [2.0, 5.0]
    x = torch.tensor([[1.0, 2.0], [3.0, 6.0]])
    mu, var = fdist.naive_sync_bn_stats(x)
    assert mu.tolist() == [2.0, 4.0] and var.tolist() == [1.0, 4.0]
    assert fdist.coalesced_all_reduce([x])[0] is x


def _train_worker(rank, world, port, out):
    """Bucketed gradient reducer + differentiable sync-BN statistics on gloo: every rank ends with the rank-mean gradient."""
    import torch.nn as nn

    from fullysparsefusion_b200.train import BucketedReducer

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(6, 8), nn.ReLU(), nn.Linear(8, 4), nn.ReLU(), nn.Linear(4, 2))
    unused = nn.Parameter(torch.ones(3))                    # a parameter that gets no gradient: reduced as zeros
    params = list(net.parameters()) + [unused]
    red = BucketedReducer(params, bucket_mb=200 / (1 << 20))   # 50 floats per bucket: several buckets
    x = torch.arange(30, dtype=torch.float32).view(5, 6) * (rank + 1) / 10
    local = []
    for step in range(2):                                   # two steps: the buckets are reusable
        for p in params:
            p.grad = None
        red.begin()
        net(x).pow(2).sum().backward()
        local = [p.grad.clone() for p in net.parameters()]
        red.finish()
    grads = [p.grad.clone() for p in net.parameters()]
    gathered = [None] * world
    dist.all_gather_object(gathered, [g.tolist() for g in local])
    # differentiable statistics: d/dx of sum(mean over ranks) = 1 / (rows * world) per element... times world (sum of all ranks' losses)
    y = (torch.ones(3, 2) * (rank + 1)).requires_grad_(True)
    mu, var = fdist.naive_sync_bn_stats(y)
    (mu.sum() + var.sum()).backward()
    if rank == 0:
        out.put(dict(n_buckets=len(red.buckets), grads=[g.tolist() for g in grads], local=gathered, unused=unused.grad.tolist(),
                     dy=y.grad.tolist(), mu=mu.tolist(), var=var.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_reducer_and_differentiable_sync_bn():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["n_buckets"] >= 2
    for i, g in enumerate(res["grads"]):                      # reduced gradient = mean of the two ranks' local gradients
        want = (torch.tensor(res["local"][0][i]) + torch.tensor(res["local"][1][i])) / 2
        assert torch.allclose(torch.tensor(g), want, rtol=1e-6, atol=1e-6)
    assert res["unused"] == [0.0, 0.0, 0.0]
    # ranks hold constant rows 1 and 2: mean 1.5, E[x^2] 2.5, var 0.25
    assert res["mu"] == [1.5, 1.5] and all(abs(v - 0.25) < 1e-6 for v in res["var"])
    # loss_r = sum(mu) + sum(var) on both ranks; d(mu)/dy = 1/(3*2); d(var)/dy = (2 y - 2 mu)/(3*2); both ranks' losses flow back
    # through the all-reduce: total = 2 * (1/6 + (2*1 - 3)/6) = 0 for rank 0's rows
    assert all(abs(v) < 1e-6 for row in res["dy"] for v in row)
