"""Data-parallel driver: images are independent, so a batch shards over the GPUs of a box with no collective
on the data path — each rank (one process per GPU) codes its contiguous block and rank 0 gathers the python
results (bit streams / metrics) on the host (the reference's only analogue is nn.DataParallel
for training, train.py:270-271)."""
from __future__ import annotations

from typing import Any, Callable, List, Optional, Sequence, Tuple

import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; the first n_items % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def run_sharded(fn: Callable[[Any], Any], items: Sequence[Any], group: Optional[dist.ProcessGroup] = None,
                dst: int = 0) -> Optional[List[Any]]:
    """Apply `fn` to this rank's block of `items`; returns the results of ALL items in order on rank `dst`
    (None elsewhere).  Without an initialised process group it degenerates to a plain map."""
    if not (dist.is_available() and dist.is_initialized()):
        return [fn(it) for it in items]
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_bounds(len(items), rank, world)
    mine = [fn(items[i]) for i in range(lo, hi)]
    gathered: Optional[List[Any]] = [None] * world if rank == dst else None
    dist.gather_object(mine, gathered, dst=dst, group=group)
    if rank != dst:
        return None
    out: List[Any] = []
    for part in gathered:
        out.extend(part)
    return out
