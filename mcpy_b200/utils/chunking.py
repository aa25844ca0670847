"""Replica-batch partitioning (mirror of ``mcpy/utils/chunking.py:11-26``).

Same contract as the reference's ``chunk_ranges`` -- ``None`` / non-positive /
oversized ``chunk_size`` means one full range, ``n_items == 0`` yields nothing --
plus :func:`shard_range`, the block partition of replicas over ranks that
SURVEY.md section 8e prescribes (rank g owns ``[g*R/G, (g+1)*R/G)``).
"""
from __future__ import annotations

from typing import Iterator, Optional, Tuple


def chunk_ranges(n_items: int, chunk_size: Optional[int]) -> Iterator[Tuple[int, int]]:
    if n_items <= 0:
        return
    if chunk_size is None or chunk_size <= 0 or chunk_size >= n_items:
        yield (0, n_items)
        return
    start = 0
    while start < n_items:
        stop = min(start + chunk_size, n_items)
        yield (start, stop)
        start = stop


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block of ``range(n_items)`` owned by ``rank``; the first
    ``n_items % world_size`` ranks get one extra item."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f'bad rank/world_size: {rank}/{world_size}')
    base, extra = divmod(max(n_items, 0), world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)
