"""GPU diagnostic: locate the worst-parity cells of an optically thick uniform grid (both arithmetic modes)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import radiativetransfer_b200 as rt  # noqa: E402
from oracle import ftte_oracle as fo  # noqa: E402
from radiativetransfer_b200 import workloads as W  # noqa: E402

bg = W.uvb_background(3.0)
n = 24
g = W.uniform_grid(n, seed=12, tau_lo=1e-3, tau_hi=300.0)
og = fo.OracleGrid(n, g["level"], g["HI"], g["HeI"], g["HeII"], box_size=g["box_size"])
o = og.diffuse(bg["uvb"], bg["beta"])
t = rt.Transport(0)
t.set_grid(n, g["level"], g["HI"], g["HeI"], g["HeII"], g["rho"], g["abun2"], g["box_size"])
res = {}
for name, mode in (("faithful", rt.MATH_FAITHFUL), ("fast", rt.MATH_FAST)):
    t.set_math(mode)
    J, _ = t.diffuse(bg["uvb"], bg["beta"])
    res[name] = J.copy()
    rel = np.abs(J - o["J"]) / np.maximum(np.abs(o["J"]), 1e-300)
    idx = np.argsort(rel.ravel())[::-1][:8]
    print(name, "worst cells:")
    for ii in idx:
        gi, c = np.unravel_index(ii, rel.shape)
        x, y, z = c // (n * n), (c // n) % n, c % n
        print(f"  g={gi} cell=({x},{y},{z}) J_gpu={J[gi, c]:.17e} J_cpu={o['J'][gi, c]:.17e} rel={rel[gi, c]:.3e}")
    # per-direction contribution of the worst cell
    gi, c = np.unravel_index(idx[0], rel.shape)
    worst = []
    for r in range(192):
        Jr, _ = t.diffuse(bg["uvb"], bg["beta"], rays=np.array([r], dtype=np.int32))
        orr = og.diffuse(bg["uvb"], bg["beta"], ray_begin=r, ray_end=r + 1)["J"]
        d = abs(Jr[gi, c] - orr[gi, c])
        worst.append((d, r, Jr[gi, c], orr[gi, c]))
    worst.sort(reverse=True)
    for d, r, a, b in worst[:4]:
        print(f"    ray {r}: gpu {a:.17e} cpu {b:.17e} diff {d:.3e}")
print("fast vs faithful max rel", np.max(np.abs(res["fast"] - res["faithful"]) / np.maximum(res["faithful"], 1e-300)))
