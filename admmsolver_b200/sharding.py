"""Batch sharding across the GPUs of one box: problems are independent, so the batch
index is partitioned into contiguous slabs, one per rank; shared operators are replicated.  The only
collective of the path is the optional all-reduce of the squared-norm partials (batch-wide
stopping criterion), issued by ``batch.SharedSpM`` through ``torch.distributed``."""
from __future__ import annotations

from typing import Tuple


def shard_range(nb: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slab [begin, end) of rank ``rank``; slab sizes differ by at most one."""
    assert 0 <= rank < world and nb >= 0
    base, rem = divmod(nb, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)
