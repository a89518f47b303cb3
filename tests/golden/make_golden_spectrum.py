"""Golden fixture for the escaping spectrum ndotSpectrum(300) / ndotDust / highestPixelLevel (equiSources.f90:3206-3224,
:3316): libm oracle, 12^3 grid in a 400 kpc box so that the rays cross the last output radius (100 kpc)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ftte_oracle  # noqa: E402
from radiativetransfer_b200 import workloads as W  # noqa: E402

here = os.path.dirname(os.path.abspath(__file__))
sp = W.synthetic_spectra()
S24 = float(np.float32(6.3e-18))
g = W.uniform_grid(12, seed=41, box_kpc=400.0, tau_lo=1e-3, tau_hi=0.3, beta24=S24)
og = ftte_oracle.OracleGrid(g["nx"], g["level"], g["HI"], g["HeI"], g["HeII"], g["rho"], g["abun2"], g["box_size"])
src, wt = np.array([(6 * 12 + 5) * 12 + 6, (5 * 12 + 6) * 12 + 7], dtype=np.int32), np.array([1, 2], dtype=np.int32)
o = og.point(sp, src, wt, dust_approximation=1)
assert o["status"] == 0 and o["ndot_spectrum"].min() > 0
np.savez_compressed(os.path.join(here, "point_spectrum_12.npz"), seed=41, box_kpc=400.0, src=src, wt=wt, dust=1,
                    ndot_spectrum=o["ndot_spectrum"], ndot_dust=o["ndot_dust"], ndot_remaining=o["ndot_remaining"],
                    ndot_boundary=o["ndot_boundary"], highest_pixel_level=o["highest_pixel_level"], nseg=o["nseg"])
print("wrote point_spectrum_12.npz", o["nseg"], o["highest_pixel_level"], o["ndot_spectrum"][:, ::100])
