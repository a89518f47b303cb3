"""GPU experiment: point-source pass with the three deposition modes (RED.ADD, sort every pass, planned)"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import radiativetransfer_b200 as rt
import bench
n, g, src, wt, sp = bench.point_inputs(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "point-128^3-amr-100src"])
N = g["level"].size
t = rt.Transport(device=0)
t.set_grid(g["nx"], g["level"], g["HI"], g["HeI"], g["HeII"], g["rho"], g["abun2"], g["box_size"])
R = torch.zeros(6, N, dtype=torch.float64, device="cuda:0")
s = torch.cuda.current_stream().cuda_stream
ref = None
for mode in (0, 2, 1, 0, 2):
    t.set_tuning(point_deposit=mode)
    ms = []
    for rep in range(5):
        R.zero_()
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        nseg = t.point_device(sp, src, wt, R.data_ptr(), stream=s)
        torch.cuda.synchronize()
        ms.append((time.perf_counter() - w0) * 1e3)
    Rh = R.cpu().numpy()
    if ref is None:
        ref = Rh
    m = ref != 0
    print(f"deposit mode {mode}: wall ms {['%.1f' % x for x in ms]} device_ms(last) {t.last_stats()['device_ms']:.2f} nseg {nseg} "
          f"max rel diff to mode 0: {np.max(np.abs(Rh[m] - ref[m]) / np.abs(ref[m])):.2e}", flush=True)
t.close()
