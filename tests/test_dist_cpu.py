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
    if rank == 0:
        out.put((gathered, t.tolist(), [fdist.frame_seed(r, i) for r in range(world) for i in range(3)]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_max_timing():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, t, seeds = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(sum(gathered, [])) == list(range(7))            # a partition of the frames
    assert gathered[0] == [0, 2, 4, 6] and gathered[1] == [1, 3, 5]
    assert t == [11.0, 5.0]                                        # MAX over ranks
    assert len(set(seeds)) == len(seeds)                           # disjoint synthetic streams
    assert fdist.throughput(10, 2, 0.5) == 40.0


def test_single_process_is_identity():
    t = torch.tensor([3.0])
    assert fdist.max_over_ranks(t).item() == 3.0
    assert fdist.frame_ids(0, 1, 3) == [0, 1, 2]
