"""GPU experiment driver for the point-source path: times rtb200_point_device on a config-3 style workload
(n^3 base grid, one refined level over the central (n/4)^3 base cells, sources inside the refined region)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import radiativetransfer_b200 as rt  # noqa: E402
from radiativetransfer_b200 import workloads as W  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=128)
ap.add_argument("--nsrc", type=int, default=100)
ap.add_argument("--modes", nargs="+", default=["fast", "faithful"])
ap.add_argument("--dust", type=int, default=0)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--uniform", action="store_true")
ap.add_argument("--deposit", type=int, default=0)
ap.add_argument("--min-blocks", type=int, nargs="+", default=[4])
args = ap.parse_args()

t0 = time.time()
g, src = W.point_workload(args.n, args.nsrc, uniform=args.uniform)
print(f"grid: {g['level'].size} leaves, {args.nsrc} sources, built in {time.time() - t0:.1f} s", flush=True)
sp = W.synthetic_spectra()
t = rt.Transport(device=0)
t.set_grid(g["nx"], g["level"], g["HI"], g["HeI"], g["HeII"], g["rho"], g["abun2"], g["box_size"])
t.set_tuning(point_deposit=args.deposit)
R = torch.zeros(6, g["level"].size, dtype=torch.float64, device="cuda:0")
s = torch.cuda.current_stream().cuda_stream
wt = np.ones(args.nsrc, dtype=np.int32)
for mode, minb in [(m, b) for m in args.modes for b in (args.min_blocks if m == "fast" else [4])]:
    t.set_math(rt.MATH_FAST if mode == "fast" else rt.MATH_FAITHFUL)
    t.set_tuning(point_min_blocks=minb)
    for rep in range(args.reps + 1):
        R.zero_()
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        nseg = t.point_device(sp, src, wt, R.data_ptr(), dust_approximation=args.dust, stream=s)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - w0) * 1e3
        st = t.last_stats()
        print(f"n={args.n} nsrc={args.nsrc} dust={args.dust} deposit={args.deposit} mode={mode} min_blocks={minb}: device_ms={st['device_ms']:.2f} wall_ms={wall:.2f} "
              f"nseg={nseg} seg/s={nseg / st['device_ms'] * 1e3:.3e} alg GB/s={st['algorithmic_bytes'] / st['device_ms'] / 1e6:.1f} "
              f"sum(krate24)={float(R[0].sum()):.6e}", flush=True)
t.close()
