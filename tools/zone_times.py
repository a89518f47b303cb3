"""times the uniform sweep zone by zone (1 GPU): input for the direction-sharding cost model"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import radiativetransfer_b200 as rt
from radiativetransfer_b200 import sharding, workloads as W
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
bg = W.uvb_background(3.0)
g = W.uniform_grid(n, seed=1)
t = rt.Transport(device=0)
t.set_grid(g["nx"], g["level"], g["HI"], g["HeI"], g["HeII"], g["rho"], g["abun2"], g["box_size"])
J = torch.zeros(3, n ** 3, dtype=torch.float64, device="cuda:0")
s = torch.cuda.current_stream().cuda_stream
zone, cost = sharding.direction_costs(3, n)
for z in range(1, 25):
    rays = np.where(zone == z)[0].astype(np.int32)
    for rep in range(3):
        t.diffuse_device(bg["uvb"], bg["beta"], J.data_ptr(), rays=rays, stream=s)
        torch.cuda.synchronize()
        ms = t.last_stats()["sweep_ms"]
    print(f"zone {z:2d} ndir {rays.size:2d} segs/col {cost[rays].sum():8.0f} sweep_ms {ms:7.3f} us/seg-col {ms * 1e3 / cost[rays].sum():.4f}", flush=True)
t.close()
