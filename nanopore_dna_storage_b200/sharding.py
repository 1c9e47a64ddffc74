"""Read sharding across the GPUs of one box: reads are independent, so each rank
decodes its own subset and the lists are gathered on the host -- no data-path collective.

`plan` deals reads to ranks longest-first onto the least-loaded rank (work ~ number of posterior
blocks), which keeps per-rank step counts balanced; `decode_sharded` runs a rank-local decode
function and gathers the per-read results on rank 0 with torch.distributed's object gather (the
result of a read is a few hundred bytes).
"""
from __future__ import annotations

from typing import Callable, List, Sequence


def plan(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Indices of the reads each rank decodes."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    load = [0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        shards[r].append(i)
        load[r] += int(lengths[i])
    for s in shards:
        s.sort()
    return shards


def decode_sharded(posts: Sequence, decode_fn: Callable[[list], list], rank: int, world_size: int,
                   gather: bool = True):
    """decode_fn(list of posteriors) -> list of per-read results.  Returns the full result list in
    input order on rank 0 (None elsewhere) when gather=True, else this rank's {index: result}."""
    shards = plan([len(p) for p in posts], world_size)
    mine = shards[rank]
    local = decode_fn([posts[i] for i in mine]) if mine else []
    local_map = dict(zip(mine, local))
    if not gather or world_size == 1:
        return [local_map[i] for i in range(len(posts))] if world_size == 1 else local_map
    import torch.distributed as dist
    gathered = [None] * world_size if rank == 0 else None
    dist.gather_object(local_map, gathered, dst=0)
    if rank != 0:
        return None
    merged = {}
    for m in gathered:
        merged.update(m)
    return [merged[i] for i in range(len(posts))]
