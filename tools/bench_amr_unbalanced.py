"""GPU experiment: nested-grid sweep on a grid that violates the 2:1 balance (a level-3 box inside level-0 cells:
per-leaf `done` flags + deferred list, one launch per wave) next to a balanced grid of similar size"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import radiativetransfer_b200 as rt
from radiativetransfer_b200 import workloads as W
bg = W.uvb_background(3.0)
cases = (("unbalanced: 64^3, box [0.4,0.6)^3 refined 3 levels at once", W.nested_grid(64, 3, W.central_box_refine(0.4, 0.6, levels=3), seed=5)),
         ("balanced: 64^3 + 3 levels around a disc", W.nested_grid(64, 3, W.disc_refine(3, r0=0.5, h0=0.12), seed=5)))
for name, g in cases:
    N = g["level"].size
    t = rt.Transport(device=0)
    t.set_grid(g["nx"], g["level"], g["HI"], g["HeI"], g["HeII"], g["rho"], g["abun2"], g["box_size"])
    J = torch.zeros(3, N, dtype=torch.float64, device="cuda:0")
    s = torch.cuda.current_stream().cuda_stream
    ms = []
    for rep in range(4):
        nseg = t.diffuse_device(bg["uvb"], bg["beta"], J.data_ptr(), stream=s)
        torch.cuda.synchronize()
        st = t.last_stats()
        ms.append(st["device_ms"])
    a = J.clone()
    t.diffuse_device(bg["uvb"], bg["beta"], J.data_ptr(), stream=s)
    torch.cuda.synchronize()
    print(f"{name}: {N} leaves, levels {np.bincount(g['level']).tolist()}, nseg {nseg}, ms {['%.2f' % m for m in ms]}, "
          f"launches {st['launches']}, {nseg / ms[-1] * 1e3:.3e} segment updates/s, run-to-run identical {bool(torch.equal(a, J))}, err {t.device_error()}", flush=True)
    t.close()
