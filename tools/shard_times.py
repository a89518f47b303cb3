"""times one rank's direction shard of an N-GPU run on a single GPU, for several directions-per-task settings"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import radiativetransfer_b200 as rt
from radiativetransfer_b200 import sharding, workloads as W
n = 256
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
bg = W.uvb_background(3.0)
g = W.uniform_grid(n, seed=1)
t = rt.Transport(device=0)
t.set_grid(g["nx"], g["level"], g["HI"], g["HeI"], g["HeII"], g["rho"], g["abun2"], g["box_size"])
J = torch.zeros(3, n ** 3, dtype=torch.float64, device="cuda:0")
s = torch.cuda.current_stream().cuda_stream
shards = sharding.shard_directions(world, n_angular_level=3, nx=n)
for rank in (0, world // 2, world - 1):
    for dpt, pdl in ((0, 0), (0, 1)):
        t.set_tuning(dirs_per_task=dpt, pdl=pdl)
        for rep in range(3):
            t.diffuse_device(bg["uvb"], bg["beta"], J.data_ptr(), rays=shards[rank], stream=s)
            torch.cuda.synchronize()
            st = t.last_stats()
        print(f"world {world} rank {rank} ndir {len(shards[rank])} dpt {dpt} pdl {pdl}: sweep_ms {st['sweep_ms']:.3f} total_ms {st['device_ms']:.3f} launches {st['launches']}", flush=True)
t.close()
