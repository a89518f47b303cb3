"""GPU experiment: nested-grid sweep as one streamed launch / one launch per wave, per-item arrays in wave order / leaf
order, thin rule on / off"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import radiativetransfer_b200 as rt
from radiativetransfer_b200 import workloads as W
bg = W.uvb_background(3.0)
for n, levels in ((64, 3), (128, 2)):
    g = W.nested_grid(n, levels, W.disc_refine(levels), seed=5)
    N = g["level"].size
    t = rt.Transport(device=0)
    t.set_grid(g["nx"], g["level"], g["HI"], g["HeI"], g["HeII"], g["rho"], g["abun2"], g["box_size"])
    J = torch.zeros(3, N, dtype=torch.float64, device="cuda:0")
    s = torch.cuda.current_stream().cuda_stream
    ref = None
    for stream, slots, thin, pdl, mb in ((0, 1, 1, 1, 6), (1, 1, 1, 1, 6), (1, 1, 1, 1, 8), (1, 1, 1, 1, 5), (0, 1, 1, 1, 8)):
        t.set_tuning(amr_stream=stream, amr_slots=slots, amr_thin=thin, pdl=pdl, amr_min_blocks=mb)
        ms = []
        for rep in range(5):
            t.diffuse_device(bg["uvb"], bg["beta"], J.data_ptr(), stream=s)
            torch.cuda.synchronize()
            ms.append(t.last_stats()["device_ms"])
        Jh = J.cpu().numpy()
        if ref is None:
            ref = Jh
        print(f"{n}^3+{levels} ({N} leaves) stream={stream} slots={slots} thin={thin} pdl={pdl} min_blocks={mb}: ms {['%.2f' % m for m in ms]} "
              f"max rel diff to first {np.max(np.abs(Jh - ref) / np.maximum(np.abs(ref), 1e-290)):.2e} err={t.device_error()}", flush=True)
    t.close()
