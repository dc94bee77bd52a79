"""How the hot path spreads over GPUs: reads are independent, so the work shards by read batch
with no data-path collective.  One process per GPU; process `rank` of `world` takes the query volumes
rank, rank + world, ... exactly like the reference's `-grid "node_id num_nodes"` loop
(src/app/map/main.c:35-55: `for (qi = node_id; qi < num_query_vols; qi += num_nodes)`), the subject .pac and
its index are replicated.  torch.distributed is used for the barrier and for reducing the timings only."""
from __future__ import annotations


def volumes_of_rank(n_volumes: int, rank: int, world: int) -> list[int]:
    """Query volumes process `rank` maps (the -grid rule)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank / world")
    return list(range(rank, n_volumes, world))


def read_seed_of_rank(seed: int, rank: int) -> int:
    """bench.py: every rank draws its own reads (weak scaling), the reference genome is shared."""
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
