"""Direction sharding for the multi-GPU diffuse sweep (every GPU holds the full grid).

Directions are independent, so the path shards with no data-path exchange; the per-leaf Jmean fields of the ranks are
summed afterwards (reduce-scatter, csrc/multi.cu).  The rule lives in the library (rtb200_shard_directions, used by the
device groups internally): whole zones are kept together where possible because the uniform-grid kernel sweeps the
directions of a zone in one pass (zone = one of the 24 index rotations, rotateIndicesModule.f90); zones are cut until
there are three pieces per rank, then assigned longest-processing-time-first on their segment counts (times an optional
per-sweep-axis cost factor).  The assignment is a pure function of (world, nAngularLevel, nx, factors): every run and
every rank computes the same shards.  This module is the Python view of it.
"""
import ctypes as C

import numpy as np

from . import _lib, solver


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


def shard_directions(world, n_angular_level=3, nx=32, zone_cost=None):
    """list (length `world`) of int32 arrays of HEALPix pixel numbers; every direction appears exactly once."""
    nrays = 12 * 4 ** (n_angular_level - 1)
    zc = None if zone_cost is None else np.ascontiguousarray(zone_cost, dtype=np.float64)
    out = []
    for rank in range(world):
        r = np.zeros(nrays, dtype=np.int32)
        n = C.c_int32(0)
        st = _lib.lib().rtb200_shard_directions(int(world), int(n_angular_level), int(nx),
                                                None if zc is None else zc.ctypes.data_as(C.c_void_p), rank,
                                                r.ctypes.data_as(C.c_void_p), nrays, C.byref(n))
        _lib.check(st, "rtb200_shard_directions")
        out.append(r[:n.value].copy())
    assert sum(s.size for s in out) == nrays
    return out
