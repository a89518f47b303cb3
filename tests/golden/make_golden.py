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
