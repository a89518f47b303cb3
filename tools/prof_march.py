"""one short run of the uniform sweep for ncu (n and variant from argv)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import radiativetransfer_b200 as rt
from radiativetransfer_b200 import workloads as W
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
march = int(sys.argv[2]) if len(sys.argv) > 2 else 1
nrays = int(sys.argv[3]) if len(sys.argv) > 3 else 24
bg = W.uvb_background(3.0)
g = W.uniform_grid(n, seed=1)
t = rt.Transport(device=0)
t.set_grid(g["nx"], g["level"], g["HI"], g["HeI"], g["HeII"], g["rho"], g["abun2"], g["box_size"])
t.set_tuning(march=march)
J = torch.zeros(3, n ** 3, dtype=torch.float64, device="cuda:0")
s = torch.cuda.current_stream().cuda_stream
for rep in range(2):
    nseg = t.diffuse_device(bg["uvb"], bg["beta"], J.data_ptr(), rays=list(range(nrays)), stream=s)
    torch.cuda.synchronize()
    print(t.last_stats())
t.close()
