"""Request-level data parallelism: utterances are independent, so a
box of N GPUs is N full replicas and a request list is partitioned by utterance.
No collective sits on the synthesis path; torch.distributed is only used by the
bench / server front end for the barrier and for gathering timings."""
from __future__ import annotations

from typing import List, Sequence, Tuple


def partition(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of ``n_items`` owned by ``rank`` (sizes differ by at most 1)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def least_loaded(loads: Sequence[float], costs: Sequence[float]) -> List[int]:
    """Greedy longest-processing-time assignment of request costs (e.g. expected tokens) to
    replicas; returns the replica index per request.  Used by the server dispatcher."""
    loads = list(loads)
    order = sorted(range(len(costs)), key=lambda i: -costs[i])
    out = [0] * len(costs)
    for i in order:
        r = min(range(len(loads)), key=lambda j: loads[j])
        out[i] = r
        loads[r] += costs[i]
    return out
