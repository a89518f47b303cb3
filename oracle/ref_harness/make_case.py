#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Writes the input files the UNMODIFIED reference driver needs for a small synthetic case
into a run directory (none of the reference's data files ship with it):

    inputParameters                         key = value lines (equiSources.f90:100-128)
    case_met.h4                             input grid, container format of hdf4_stub.c (equiSources.f90:316-423)
    case_sources.dat                        level x y z age rows (equiSources.f90:734-771)
    model4?-salpeter-burst34/spectrum.out   five starburst99-style tables (equiSources.f90:840-892)
    smc_dust_parameters.dat, lmc_...        7 x 5 extinction-fit rows (dustModule.f90:15-24)

    python make_case.py <rundir> [--case config1|nested] [--mode 8]

The grid comes from the same generators as the repository's tests (radiativetransfer_b200/workloads.py).  The driver
re-derives HI etc. from (lT, lnH, lx) in single precision and runs its initial equilibrium first, so the comparison
(compare.py) does NOT rely on these inputs: it reads the state the patched driver dumps right before the transport.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from radiativetransfer_b200 import formats as F  # noqa: E402
from radiativetransfer_b200 import workloads as W  # noqa: E402


def level_lists(g, box_kpc):
    """per-level cell lists (every node of the octree, internal nodes included) from leaf arrays in pre-order"""
    n = g["nx"]
    level = g["level"].astype(int)
    cx, cy, cz = F.leaf_centres(n, level)
    nH = g["rho"] * W.PSI / W.MP
    vals = dict(lT=np.full(level.size, 4.0), lnH=np.log10(nH), lx=np.log10(np.minimum(g["HI"] / nH, 1.0)),
                ab=g["abun2"])
    lmax = level.max()
    out = []
    for L in range(lmax + 1):
        # nodes of depth L = leaves of level L + ancestors (at depth L) of deeper leaves; an ancestor takes the values of
        # its first descendant leaf (the driver overwrites them with the children anyway)
        size = 1.0 / (n << L)
        ix = np.floor(cx / size).astype(np.int64); iy = np.floor(cy / size).astype(np.int64); iz = np.floor(cz / size).astype(np.int64)
        m = level >= L
        key = (ix[m] * (n << L) + iy[m]) * (n << L) + iz[m]
        _, first = np.unique(key, return_index=True)
        sel = np.where(m)[0][first]
        pos = np.stack([(ix[sel] + 0.5) * size, (iy[sel] + 0.5) * size, (iz[sel] + 0.5) * size], axis=1) * box_kpc
        ab = np.zeros((sel.size, 4)); ab[:, 1] = vals["ab"][sel]
        out.append(dict(pos=pos, lT=vals["lT"][sel], lnH=vals["lnH"][sel], lx=vals["lx"][sel], abun=ab))
    return out


def write_spectrum_file(path, wl_A, lum_two_slices):
    """37 time slices of 1221 wavelengths in the fixed columns the driver parses (cols 2-13 time [yr], 14-28 wavelength
    [A], 29-41 log10 luminosity); slices 9 and 10 carry the two synthetic spectra, so that timeReadTable = 10 Myr
    (equiSources.f90:1236) brackets them"""
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        f.write(" MODEL DESIGNATION: rtb200 synthetic\n")
        for it in range(37):
            t = 1.0e6 * (it + 1)
            lum = lum_two_slices[0] if it < 9 else lum_two_slices[1]
            f.write(" TIME [YR]    WAVELENGTH [A]   LOG TOTAL  \n")
            f.write(" \n")
            f.write(" \n")
            for w, l in zip(wl_A, lum):
                f.write(" %12.5E%15.5f%13.5f\n" % (t, w, l))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rundir")
    ap.add_argument("--case", default="config1", choices=["config1", "nested"])
    ap.add_argument("--mode", type=int, default=8)       # 8 = stellar + UVB transfer (equiSources.f90:65-67)
    a = ap.parse_args()
    os.makedirs(a.rundir, exist_ok=True)
    box_kpc = 100.0
    s24 = float(np.float32(6.3e-18))
    if a.case == "config1":   # BASELINE config 1: 32^3, one source in cell (17,17,17), H-dominated
        g = W.uniform_grid(32, seed=20240601, box_kpc=box_kpc, tau_lo=1e-3, tau_hi=1.0, beta24=s24, helium=False)
        src = [(16.5 / 32 * box_kpc,) * 3]
    else:
        g = W.nested_grid(8, 2, W.central_box_refine(0.3, 0.7, levels=2), seed=32, box_kpc=box_kpc, tau_lo=1e-3, tau_hi=0.5, beta24=s24)
        src = [(0.52 * box_kpc, 0.47 * box_kpc, 0.5 * box_kpc), (0.31 * box_kpc, 0.64 * box_kpc, 0.55 * box_kpc)]
    F.write_grid_container(os.path.join(a.rundir, "case_met.h4"), level_lists(g, box_kpc), metals=True)
    with open(os.path.join(a.rundir, "case_sources.dat"), "w") as f:
        for x, y, z in src:
            f.write(f"1 {x:.6f} {y:.6f} {z:.6f} 5.0\n")
    sp = W.synthetic_spectra()
    wl_A = sp["wavelength"] * 1.0e8
    for m in range(5):
        write_spectrum_file(os.path.join(a.rundir, f"model4{m + 1}-salpeter-burst34", "spectrum.out"), wl_A, sp["lum"][m])
    for name in ("smc_dust_parameters.dat", "lmc_dust_parameters.dat"):
        np.savetxt(os.path.join(a.rundir, name), sp["a_dust"], fmt="%.8e")
    with open(os.path.join(a.rundir, "inputParameters"), "w") as f:
        f.write("dustApproximation = 0\nselfShieldingThreshold = 1.\ncurrentRedshift = 3.\nmassStellarParticle = 1\n"
                "grid = 'case_met'\nsources = 'case_sources.dat'\n"
                f"mode = {a.mode}\nupperAgeLimit = 10.\nsphDir = './'\nsynthesisDir = './'\nrestart = 0\nuvbCoefficient = 1.\n")
    print("wrote case", a.case, "with", g["level"].size, "leaves into", a.rundir)


if __name__ == "__main__":
    main()
