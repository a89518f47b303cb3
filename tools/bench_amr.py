"""GPU experiment driver: diffuse sweep on a nested (AMR) grid, config-5 style (64^3 base + 3 levels around a disc)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import radiativetransfer_b200 as rt
from radiativetransfer_b200 import workloads as W
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
levels = int(sys.argv[2]) if len(sys.argv) > 2 else 3
nrays = int(sys.argv[3]) if len(sys.argv) > 3 else 192
bg = W.uvb_background(3.0)
t0 = time.time()
g = W.nested_grid(n, levels, W.disc_refine(levels), seed=5)
N = g["level"].size
print(f"grid {n}^3 + {levels} levels: {N} leaves, levels histogram {np.bincount(g['level']).tolist()}, built in {time.time()-t0:.1f}s", flush=True)
t = rt.Transport(device=0)
t.set_grid(g["nx"], g["level"], g["HI"], g["HeI"], g["HeII"], g["rho"], g["abun2"], g["box_size"])
J = torch.zeros(3, N, dtype=torch.float64, device="cuda:0")
s = torch.cuda.current_stream().cuda_stream
for rep in range(3):
    w0 = time.perf_counter()
    nseg = t.diffuse_device(bg["uvb"], bg["beta"], J.data_ptr(), rays=None if nrays == 192 else list(range(nrays)), stream=s)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - w0) * 1e3
    st = t.last_stats()
    print(f"rep {rep}: device_ms={st['device_ms']:.2f} wall_ms={wall:.1f} nseg={nseg} seg/s={nseg/st['device_ms']*1e3:.3e} "
          f"alg GB/s={st['algorithmic_bytes']/st['device_ms']/1e6:.1f} launches={st['launches']} err={t.device_error()}", flush=True)
t.close()
