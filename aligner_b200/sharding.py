"""Utterance sharding across GPUs.

Utterances are independent (reference: core.pyx:44-45, the batch loop has no cross-item state), so a batch is
partitioned across ranks with no collective on the data path.  For mixed-length batches the cost of an item is
its number of DP cells ``t_x * t_y``; balancing by count can leave one GPU with all the long utterances.
``balance_shards`` is the classic longest-processing-time-first greedy: items in descending cost order, each to
the least-loaded rank.  ``lpt_order`` is the same idea inside one GPU: the persistent grid takes work items in
index order, so handing it the longest utterances first shortens the tail of a ragged batch.
"""
from __future__ import annotations

import heapq
from typing import List, Sequence

import numpy as np

__all__ = ["item_cost", "balance_shards", "lpt_order", "shard_loads"]


def item_cost(t_x: Sequence[int], t_y: Sequence[int]) -> np.ndarray:
    """DP cells per utterance (the unit of BASELINE.json's metric)."""
    return np.asarray(t_x, dtype=np.int64) * np.asarray(t_y, dtype=np.int64)


def balance_shards(t_x: Sequence[int], t_y: Sequence[int], world_size: int) -> List[np.ndarray]:
    """Indices of the utterances each rank aligns; every index appears exactly once.

    Deterministic (ties broken by index, then by rank), so every rank can compute the same plan locally from the
    lengths alone -- no communication needed to agree on it.  Within a shard the indices are in descending cost
    order (see ``lpt_order``).
    """
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    cost = item_cost(t_x, t_y)
    order = np.lexsort((np.arange(cost.size), -cost))          # descending cost, stable in index
    heap = [(0, r) for r in range(world_size)]
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        load, r = heapq.heappop(heap)
        shards[r].append(int(i))
        heapq.heappush(heap, (load + int(cost[i]), r))
    return [np.asarray(s, dtype=np.int64) for s in shards]


def lpt_order(t_x: Sequence[int], t_y: Sequence[int]) -> np.ndarray:
    """Permutation that puts the most expensive utterances first (stable)."""
    cost = item_cost(t_x, t_y)
    return np.lexsort((np.arange(cost.size), -cost)).astype(np.int64)


def shard_loads(t_x: Sequence[int], t_y: Sequence[int], shards: Sequence[np.ndarray]) -> np.ndarray:
    cost = item_cost(t_x, t_y)
    return np.asarray([int(cost[s].sum()) for s in shards], dtype=np.int64)
