"""How the hot path spreads over GPUs: reads are independent, so the work shards by read batch
with no data-path collective.  One process per GPU; process `rank` of `world` takes the query volumes
rank, rank + world, ... exactly like the reference's `-grid "node_id num_nodes"` loop
(src/app/map/main.c:35-55: `for (qi = node_id; qi < num_query_vols; qi += num_nodes)`), the subject .pac and
its index are replicated.  torch.distributed is used for the barrier, for reducing the timings and -- when a
fixed read set is spread over the ranks (strong scaling) -- for the shared work counter (WorkQueue); no tensor of
the data path ever crosses a rank boundary.  (Inside one process the drop-in tools do the same with feeder threads:
lesv_b200/host/lesv_gpus.hpp.)"""
from __future__ import annotations


def volumes_of_rank(n_volumes: int, rank: int, world: int) -> list[int]:
    """Query volumes process `rank` maps (the -grid rule)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank / world")
    return list(range(rank, n_volumes, world))


def read_seed_of_rank(seed: int, rank: int) -> int:
    """A private read set per rank (the reference genome is shared).  bench.py used this in round 1; it now gives every rank
    the same pool of batches, rotated, because the spread between private draws alone cost 4-5 % at 8 GPUs."""
    return seed + 1000 * (rank + 1)


def reduce_stats(dev_ms: float, wall_ms: float, bases: float, cells: float, launches: float, world: int, device="cpu"):
    """Whole-job figures from per-rank ones: times are the MAX over ranks (the job ends with its slowest rank),
    work counts the SUM.  Returns (dev_ms, wall_ms, bases, cells, launches)."""
    if world <= 1:
        return dev_ms, wall_ms, bases, cells, launches
    import torch
    import torch.distributed as dist
    t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device=device)
    w = torch.tensor([bases, cells, launches], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(w, op=dist.ReduceOp.SUM)
    return (*[float(x) for x in t.tolist()], *[float(x) for x in w.tolist()])


def spread(value: float, world: int, device="cpu"):
    """(max, mean, min) of a per-rank figure over the ranks: how far the slowest rank is from the average."""
    if world <= 1:
        return value, value, value
    import torch
    import torch.distributed as dist
    hi = torch.tensor([value, -value], dtype=torch.float64, device=device)
    su = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    dist.all_reduce(su, op=dist.ReduceOp.SUM)
    return float(hi[0]), float(su[0]) / world, -float(hi[1])


class WorkQueue:
    """A list of sub-batches shared by all ranks of the job: every rank calls next() until it returns None and gets
    each index at most once job-wide, in ascending order, whichever rank asks first (the slow rank of a static split
    simply takes fewer).  The counter lives in the rendezvous store of torch.distributed (an atomic add on the
    rank-0 TCP store); no tensor moves, so it works the same under nccl and gloo.  One rank: a plain counter."""

    def __init__(self, n_items: int, world: int, name: str = "lesv_queue"):
        self.n, self.world, self.key, self._i = int(n_items), int(world), name, 0
        self.store = None
        if world > 1:
            import torch.distributed as dist
            from torch.distributed import distributed_c10d as c10d
            self.store = dist.PrefixStore(name, c10d._get_default_store())

    def next(self):
        if self.store is None:
            i = self._i
            self._i += 1
        else:
            i = self.store.add(self.key, 1) - 1
        return i if i < self.n else None
