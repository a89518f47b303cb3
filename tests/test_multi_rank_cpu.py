"""world_size-2 gloo test (CPU) of the N > 1 path's host logic: direction sharding + all-reduce of the per-rank
partial Jmean fields reproduces the single-rank sum.  The per-rank compute here is the CPU oracle (test
infrastructure); on GPUs bench.py runs the same plumbing with the CUDA path and NCCL."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    from oracle import ftte_oracle as fo
    from radiativetransfer_b200 import sharding, workloads as W
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bg = W.uvb_background(3.0)
    g = W.uniform_grid(8, seed=5)
    shards = sharding.shard_directions(world, nx=8)
    og = fo.OracleGrid(8, g["level"], g["HI"], g["HeI"], g["HeII"], box_size=g["box_size"])
    o = og.diffuse_mt(bg["uvb"], bg["beta"], shards[rank], nthreads=1)
    assert o["status"] == 0
    J = torch.from_numpy(o["J"].copy())
    nseg = torch.tensor([o["nseg"]], dtype=torch.int64)
    dist.all_reduce(J)
    dist.all_reduce(nseg)
    if rank == 0:
        q.put((J.numpy(), int(nseg[0])))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_sharded_sum_equals_single_rank(world, oracle, uvbg):
    import torch.multiprocessing as mp
    from radiativetransfer_b200 import workloads as W
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    J, nseg = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = W.uniform_grid(8, seed=5)
    og = oracle.OracleGrid(8, g["level"], g["HI"], g["HeI"], g["HeII"], box_size=g["box_size"])
    full = og.diffuse(uvbg["uvb"], uvbg["beta"])
    assert nseg == full["nseg"]
    assert np.allclose(J, full["J"], rtol=1e-13, atol=0)


def _point_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    from oracle import ftte_oracle as fo
    from radiativetransfer_b200 import workloads as W
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g, src = W.point_workload(8, 5, uniform=True)
    wt = np.arange(1, src.size + 1, dtype=np.int32)
    sp = W.synthetic_spectra()
    og = fo.OracleGrid(8, g["level"], g["HI"], g["HeI"], g["HeII"], g["rho"], g["abun2"], g["box_size"])
    mine = slice(rank, None, world)                     # the round-robin source sharding of bench.py
    o = og.point(sp, src[mine], wt[mine])
    assert o["status"] == 0
    R = torch.from_numpy(o["rates"].copy())
    nseg = torch.tensor([o["nseg"]], dtype=torch.int64)
    dist.all_reduce(R)
    dist.all_reduce(nseg)
    if rank == 0:
        q.put((R.numpy(), int(nseg[0])))
    dist.destroy_process_group()


def test_point_sources_sharded_over_ranks(oracle):
    """sources are independent: per-rank rate fields add up to the single-rank result (all-reduce of 6 fields)"""
    import torch.multiprocessing as mp
    from radiativetransfer_b200 import workloads as W
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_point_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    R, nseg = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g, src = W.point_workload(8, 5, uniform=True)
    wt = np.arange(1, src.size + 1, dtype=np.int32)
    og = oracle.OracleGrid(8, g["level"], g["HI"], g["HeI"], g["HeII"], g["rho"], g["abun2"], g["box_size"])
    full = og.point(W.synthetic_spectra(), src, wt)
    assert nseg == full["nseg"]
    assert np.allclose(R, full["rates"], rtol=1e-12, atol=0)
