"""Owner-computes partition of independent preconditioner units across the GPUs of one box.

Every parameter's (Q, L) -- and every LRA block -- is independent of every other one (the reference loops per parameter
with no cross-parameter state, ddp.py:112-161, and simply replicates all of the work on every rank).  Here each unit is
assigned to exactly one rank by longest-processing-time-first greedy on the algorithmic cost model of SURVEY.md 8d;
a unit's state lives only on its owner (ZeRO-1-like saving).  Throughput mode needs no collective at all; a training job
additionally exchanges the updated parameters (one broadcast / all-gather per bucket) -- see DESIGN.md.
"""
from typing import List, Sequence, Tuple


def kron_unit_cost(m: int, n: int, dense_l: bool, dense_r: bool) -> float:
    """FLOPs of one update + one apply (full-GEMM counting, min-flop order)."""
    chain = 0.0
    if dense_l:
        chain += 2.0 * m * m * n + 2.0 * m * m * min(m, n)
    if dense_r:
        chain += 2.0 * m * n * n + 2.0 * n * n * min(m, n)
    upd = chain
    if dense_l:
        upd += 2.0 * m * m * n + 6.0 * m ** 3
    if dense_r:
        upd += 2.0 * m * n * n + 6.0 * n ** 3
    # bandwidth-bound part expressed in flop-equivalents (about 40 HBM passes over m*n at ~250 flop/byte machine balance)
    return upd + chain + 40.0 * m * n * 2 * 250.0


def lra_unit_cost(n: int, r: int, elem_bytes: int = 2) -> float:
    return (10.0 * n * r * elem_bytes + 12.0 * n * elem_bytes) * 250.0 + 14.0 * n * r * r


def lpt_partition(costs: Sequence[float], world_size: int, initial_loads: Sequence[float] | None = None) -> List[List[int]]:
    """Greedy LPT: returns, per rank, the (sorted) list of unit indices it owns. Deterministic on every rank.
    `initial_loads`: work every rank already carries (e.g. its row shard of a sharded LRA preconditioner)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = list(initial_loads) if initial_loads is not None else [0.0] * world_size
    assert len(loads) == world_size
    owned: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        owned[r].append(i)
        loads[r] += costs[i]
    return [sorted(o) for o in owned]


def owner_of(costs: Sequence[float], world_size: int) -> List[int]:
    own = [0] * len(costs)
    for r, idxs in enumerate(lpt_partition(costs, world_size)):
        for i in idxs:
            own[i] = r
    return own


def imbalance(costs: Sequence[float], parts: List[List[int]]) -> Tuple[float, float]:
    loads = [sum(costs[i] for i in p) for p in parts]
    mean = sum(loads) / len(loads)
    return max(loads) / mean if mean > 0 else 1.0, mean


def row_shard(n: int, world_size: int, rank: int, align: int = 256) -> Tuple[int, int]:
    """Rows [lo, hi) of an n-row LRA preconditioner owned by `rank`: contiguous, sizes multiples of `align` (whole bulk-copy blocks of the
    sweep kernels) except for the last rank, which takes the remainder."""
    per = -(-n // world_size)
    per = -(-per // align) * align
    lo = min(n, rank * per)
    hi = min(n, lo + per) if rank < world_size - 1 else n
    return lo, max(lo, hi)
