"""Generates the committed golden fixtures from the CPU oracle (the Fortran reference cannot be built in this image:
no Fortran compiler, no HDF4 -- so these pin the GPU path to the oracle, not the oracle to the reference)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ftte_oracle  # noqa: E402
from radiativetransfer_b200 import workloads as W  # noqa: E402

here = os.path.dirname(os.path.abspath(__file__))
bg = W.uvb_background(3.0)
g = W.uniform_grid(12, seed=2024)
og = ftte_oracle.OracleGrid(g["nx"], g["level"], g["HI"], g["HeI"], g["HeII"], box_size=g["box_size"])
o = og.diffuse(bg["uvb"], bg["beta"])
assert o["status"] == 0
np.savez_compressed(os.path.join(here, "diffuse_uniform_12.npz"), HI=g["HI"], uvb=bg["uvb"], beta=bg["beta"],
                    J=o["J"], nseg=o["nseg"])
print("wrote diffuse_uniform_12.npz", o["nseg"])

# ---- point sources: 10^3 uniform grid with helium and one refined 6^3 nested grid, two sources each, libm oracle ----
sp = W.synthetic_spectra()
S24 = float(np.float32(6.3e-18))
g = W.uniform_grid(10, seed=77, tau_lo=1e-2, tau_hi=1.0, beta24=S24)
g["abun2"] = np.random.default_rng(7).uniform(1e-3, 4e-2, 1000)
og = ftte_oracle.OracleGrid(g["nx"], g["level"], g["HI"], g["HeI"], g["HeII"], g["rho"], g["abun2"], g["box_size"])
src, wt = np.array([555, 17], dtype=np.int32), np.array([2, 1], dtype=np.int32)
o = og.point(sp, src, wt, dust_approximation=1)
assert o["status"] == 0
np.savez_compressed(os.path.join(here, "point_uniform_10.npz"), seed=77, src=src, wt=wt, dust=1, rates=o["rates"],
                    ndot_remaining=o["ndot_remaining"], ndot_boundary=o["ndot_boundary"], ndot_dust=o["ndot_dust"],
                    ndot_spectrum=o["ndot_spectrum"], nseg=o["nseg"], abun2=g["abun2"])
print("wrote point_uniform_10.npz", o["nseg"])
g = W.nested_grid(6, 2, W.central_box_refine(0.2, 0.8, levels=2), seed=78, tau_lo=1e-2, tau_hi=0.5, beta24=S24)
og = ftte_oracle.OracleGrid(g["nx"], g["level"], g["HI"], g["HeI"], g["HeII"], g["rho"], g["abun2"], g["box_size"])
cx, cy, cz = g["centres"]
src = np.array([int(np.argmin((cx - 0.5) ** 2 + (cy - 0.45) ** 2 + (cz - 0.55) ** 2)), 3], dtype=np.int32)
wt = np.array([1, 2], dtype=np.int32)
o = og.point(sp, src, wt, trace_cap=2000000)
assert o["status"] == 0
n0 = og.point(sp, src[:1], wt[:1])["nseg"]          # the trace lists source 0's segments first (depth-first per source)
np.savez_compressed(os.path.join(here, "point_amr_6.npz"), seed=78, src=src, wt=wt, dust=0, rates=o["rates"],
                    ndot_remaining=o["ndot_remaining"], ndot_boundary=o["ndot_boundary"], nseg=o["nseg"],
                    level=g["level"], trace=o["trace"], nseg_source0=n0)
print("wrote point_amr_6.npz", o["nseg"], g["level"].size)
