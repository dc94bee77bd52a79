"""Multi-GPU host logic on CPU: two processes over gloo (no GPU).  The path shards by read batch with no
data-path collective; what is distributed is the volume assignment (the reference's -grid rule)
and the reduction of the per-rank figures bench.py reports."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lesv_b200 import shard


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_vol, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.volumes_of_rank(n_vol, rank, world)
    # every rank learns every assignment: disjoint and complete
    flags = torch.zeros(n_vol, dtype=torch.int64)
    flags[mine] = 1
    dist.all_reduce(flags, op=dist.ReduceOp.SUM)
    assert bool((flags == 1).all()), flags
    # per-rank figures -> whole-job figures
    dev, wall, bases, cells, launches = shard.reduce_stats(100.0 + 50 * rank, 120.0 + 10 * rank, 1e6 * (rank + 1),
                                                           2e9 * (rank + 1), 10 + rank, world)
    assert dev == 100.0 + 50 * (world - 1) and wall == 120.0 + 10 * (world - 1)
    assert bases == 1e6 * world * (world + 1) / 2 and cells == 2e9 * world * (world + 1) / 2
    assert launches == sum(10 + r for r in range(world))
    assert shard.read_seed_of_rank(7, rank) != shard.read_seed_of_rank(7, (rank + 1) % world)
    hi, mean, lo = shard.spread(10.0 + 4 * rank, world)
    assert (hi, mean, lo) == (10.0 + 4 * (world - 1), 10.0 + 2 * (world - 1), 10.0)
    # the shared work queue: every sub-batch exactly once over both ranks, whoever asks first
    q = shard.WorkQueue(23, world)
    got = []
    while True:
        i = q.next()
        if i is None:
            break
        got.append(i)
    assert got == sorted(got)
    seen = torch.zeros(23, dtype=torch.int64)
    seen[got] = 1
    dist.all_reduce(seen, op=dist.ReduceOp.SUM)
    assert bool((seen == 1).all()), seen
    dist.barrier()
    if rank == 0:
        out.put(("ok", mine))
    dist.destroy_process_group()


def test_two_ranks_over_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 7, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    status, mine = out.get(timeout=5)
    assert status == "ok" and mine == [0, 2, 4, 6]


def test_grid_rule_matches_driver_loop():
    # qx2map.cpp: for (q = node_id; q < nqv; q += num_nodes)
    assert shard.volumes_of_rank(5, 1, 2) == [1, 3]
    assert shard.volumes_of_rank(2, 3, 4) == []
    assert shard.reduce_stats(1.0, 2.0, 3.0, 4.0, 5.0, 1) == (1.0, 2.0, 3.0, 4.0, 5.0)
    q = shard.WorkQueue(3, 1)
    assert [q.next() for _ in range(5)] == [0, 1, 2, None, None]
