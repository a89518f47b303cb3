"""Direction sharding for the multi-GPU diffuse sweep (one process per GPU, every GPU holds the full grid).

Directions are independent, so the path shards with no data-path exchange; the per-leaf Jmean fields of the ranks
are summed afterwards with one all-reduce.  Whole zones are kept together where possible because the uniform-grid
kernel sweeps the directions of a zone in one pass (zone = one of the 24 index rotations,
rotateIndicesModule.f90); the assignment is deterministic (longest-processing-time first on the exact segment
counts), so every run sums the same partial results.
"""
import numpy as np

from . import solver


def direction_costs(n_angular_level=3, nx=32):
    """izone[nrays], cost[nrays] with cost = segments per base-cell column (sum over layers of 1..3)."""
    nrays = 12 * 4 ** (n_angular_level - 1)
    zone = np.zeros(nrays, dtype=np.int32)
    cost = np.zeros(nrays)
    for r in range(nrays):
        zone[r] = solver.direction(n_angular_level, r)[0]
        p = solver.patterns(n_angular_level, r, nx)
        cost[r] = np.sum(1 + (p[:, 5] != 0) + (p[:, 8] != 0))     # xy + active xz (len != 0) + active yz (len != 0)
    return zone, cost


def shard_directions(world, n_angular_level=3, nx=32, zone=None, cost=None):
    """list (length `world`) of int32 arrays of HEALPix pixel numbers; every direction appears exactly once."""
    if zone is None or cost is None:
        zone, cost = direction_costs(n_angular_level, nx)
    nrays = zone.size
    groups = [np.where(zone == z)[0] for z in range(1, 25)]
    groups = [g for g in groups if g.size]
    # split zones until there are at least 3 pieces per rank, so LPT can balance
    while len(groups) < 3 * world and max(g.size for g in groups) > 1:
        groups.sort(key=lambda g: -cost[g].sum())
        g = groups.pop(0)
        groups += [g[: g.size // 2], g[g.size // 2:]]
    groups.sort(key=lambda g: (-cost[g].sum(), int(g[0])))
    load = np.zeros(world)
    out = [[] for _ in range(world)]
    for g in groups:
        r = int(np.argmin(load))
        out[r].append(g)
        load[r] += cost[g].sum()
    shards = [np.sort(np.concatenate(o)).astype(np.int32) if o else np.zeros(0, dtype=np.int32) for o in out]
    assert sum(s.size for s in shards) == nrays
    return shards
