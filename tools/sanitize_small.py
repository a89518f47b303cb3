"""small end-to-end calls of every path for compute-sanitizer (memcheck): uniform sweep (both kernels, both math modes),
AMR sweep, point sources (both deposition modes, trace)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import radiativetransfer_b200 as rt
from radiativetransfer_b200 import workloads as W
bg = W.uvb_background(3.0)
sp = W.synthetic_spectra()
t = rt.Transport(device=0)
g = W.uniform_grid(33, seed=1)
t.set_grid(g["nx"], g["level"], g["HI"], g["HeI"], g["HeII"], g["rho"], g["abun2"], g["box_size"])
for march in (0, 1):
    for mode in (rt.MATH_FAST, rt.MATH_FAITHFUL):
        t.set_math(mode); t.set_tuning(march=march)
        J, nseg = t.diffuse(bg["uvb"], bg["beta"], rays=list(range(0, 192, 5)))
        print("uniform", march, mode, nseg, float(J.sum()))
t.set_tuning(march=0)
g = W.nested_grid(6, 2, W.central_box_refine(0.2, 0.8, levels=2), seed=4, tau_lo=1e-3, tau_hi=0.5, beta24=6.3e-18)
t.set_grid(g["nx"], g["level"], g["HI"], g["HeI"], g["HeII"], g["rho"], g["abun2"], g["box_size"])
t.set_math(rt.MATH_FAST)
J, nseg = t.diffuse(bg["uvb"], bg["beta"], rays=list(range(0, 192, 7)))
print("amr", nseg, float(J.sum()), t.device_error())
for dep in (0, 1):
    for mode in (rt.MATH_FAST, rt.MATH_FAITHFUL):
        t.set_math(mode); t.set_tuning(point_deposit=dep)
        r = t.point(sp, [100, 2000], [1, 2], dust_approximation=2, max_pixel_level=4)
        print("point", dep, mode, r["nseg"], float(r["rates"].sum()))
t.set_tuning(point_deposit=0)
r = t.point(sp, [100], [1], max_pixel_level=3, trace_cap=100000)
print("trace", r["nseg"], r["trace"].size)
t.close()
