#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Compares what the patched reference driver dumped (ftte_reference_dump.bin,
ftte_reference_sources.bin; layout in patch_driver.py) with the CPU oracle fed the SAME state: this is the pin the
oracle lacks in an image without a Fortran compiler.

    python compare.py <rundir>            exit 0 when Jmean1..3 agree to 1e-9 relative, the six rate fields to 1e-9
                                          relative (+ the conditioning floor, tests/test_point_gpu.py), the escape
                                          diagnostics to 1e-11 and highestPixelLevel exactly
    python compare.py --selftest <tmpdir>  writes a dump FROM the oracle in the same layout and reads it back (checks
                                          this reader; runs without gfortran)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ftte_oracle as fo  # noqa: E402


class Reader:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        self.o = 0

    def take(self, dtype, n):
        a = np.frombuffer(self.b, dtype=dtype, count=n, offset=self.o)
        self.o += a.nbytes
        return a.copy()

    def done(self):
        return self.o == len(self.b)


def read_dump(rundir):
    r = Reader(os.path.join(rundir, "ftte_reference_dump.bin"))
    h = r.take("<i4", 10)
    assert h[0] == 20260001, "not a harness dump"
    d = dict(nx=int(h[1]), nleaf=int(h[2]), nstars=int(h[3]), nwave=int(h[4]), nmetal=int(h[5]), ispectrum=int(h[6]),
             dust=int(h[7]), max_pixel_level=int(h[8]), n_angular_level=int(h[9]))
    sc = r.take("<f8", 9)
    d.update(box=sc[0], coef_spectrum=sc[1], redshift=sc[2], uvb=sc[3:6], alpha=sc[6:9])
    beta = r.take("<f8", 9).reshape(3, 3)          # [group][beta24, beta25, beta26]
    ksi = r.take("<f8", 9).reshape(3, 3)
    d["beta"] = beta[:, [0, 2, 1]]                 # -> [group][beta24, beta26, beta25] (the order the sweep takes)
    d["ksi"] = ksi
    d["wavelength"] = r.take("<f8", d["nwave"])
    d["metallicity"] = r.take("<f8", d["nmetal"])
    d["lum"] = r.take("<f8", d["nmetal"] * 2 * d["nwave"]).reshape(d["nmetal"], 2, d["nwave"])
    d["a_dust"] = r.take("<f8", 35).reshape(7, 5)
    d["weight"] = r.take("<i4", d["nstars"])
    N = d["nleaf"]
    d["level"] = r.take("<i4", N).astype(np.int8)
    for k in ("HI", "HeI", "HeII", "rho", "abun2", "tgas"):
        d[k] = r.take("<f8", N)
    d["J"] = np.stack([r.take("<f8", N) for _ in range(3)])
    d["rates"] = np.stack([r.take("<f8", N) for _ in range(6)])     # krate24, 25, 26, crate24, 25, 26
    d["species_after"] = [r.take("<f8", N) for _ in range(3)]
    assert r.done(), "trailing bytes in the dump"
    s = Reader(os.path.join(rundir, "ftte_reference_sources.bin"))
    src = []
    while not s.done():
        ii = s.take("<i4", 4)
        v = s.take("<f8", 15)
        spec = s.take("<f8", 300)
        src.append(dict(istar=int(ii[0]), leaf=int(ii[1]), weight=int(ii[2]), highest=int(ii[3]), remaining=v[:7],
                        boundary=v[7:14], dust=v[14], spectrum=spec))
    d["sources"] = src
    return d


def rel(a, b, floor=1e-290):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


def compare(rundir):
    d = read_dump(rundir)
    og = fo.OracleGrid(d["nx"], d["level"], d["HI"], d["HeI"], d["HeII"], d["rho"], d["abun2"], d["box"])
    ok = True
    if np.any(d["J"] != 0):
        o = og.diffuse(d["uvb"], d["beta"], n_angular_level=d["n_angular_level"])
        e = rel(o["J"], d["J"])
        print(f"diffuse: status {o['status']}, Jmean1..3 rel L-inf {e:.3e}")
        ok &= o["status"] == 0 and e < 1e-9
    if d["sources"]:
        sp = dict(wavelength=d["wavelength"], lum=d["lum"], metallicity=d["metallicity"], coef_spectrum=d["coef_spectrum"],
                  a_dust=d["a_dust"])
        leaf = np.array([s["leaf"] for s in d["sources"]], dtype=np.int32)
        wt = np.array([s["weight"] for s in d["sources"]], dtype=np.int32)
        o = og.point(sp, leaf, wt, dust_approximation=d["dust"], max_pixel_level=d["max_pixel_level"])
        scale = np.abs(d["rates"]).sum(axis=1, keepdims=True)          # >= what the sources emit into each field
        within = np.all(np.abs(o["rates"] - d["rates"]) <= 1e-9 * np.abs(d["rates"]) + 2e-13 * scale)
        print(f"point: status {o['status']}, rates within 1e-9 (+ floor): {bool(within)}, strict rel L-inf "
              f"{rel(o['rates'][d['rates'] != 0], d['rates'][d['rates'] != 0]):.3e}")
        ok &= o["status"] == 0 and bool(within)
        for i, s in enumerate(d["sources"]):
            er = max(rel(o["ndot_remaining"][i], s["remaining"], 1e-300), rel(o["ndot_boundary"][i], s["boundary"], 1e-300),
                     rel(o["ndot_spectrum"][i], s["spectrum"], 1e-300), rel(o["ndot_dust"][i:i + 1], np.array([s["dust"]]), 1e-300))
            hp = int(o["highest_pixel_level"][i]) == s["highest"]
            print(f"  source {s['istar']}: diagnostics rel L-inf {er:.3e}, highestPixelLevel equal: {hp}")
            ok &= er < 1e-11 and hp
    print("PINNED" if ok else "MISMATCH")
    return 0 if ok else 1


def selftest(tmp):
    """a dump written from the oracle's own results, in the layout the patched driver writes"""
    from radiativetransfer_b200 import workloads as W
    os.makedirs(tmp, exist_ok=True)
    g = W.nested_grid(4, 1, W.central_box_refine(0.25, 0.75, levels=1), seed=3, tau_lo=1e-2, tau_hi=0.5,
                      beta24=float(np.float32(6.3e-18)))
    bg = W.uvb_background(3.0)
    sp = W.synthetic_spectra()
    N = g["level"].size
    og = fo.OracleGrid(4, g["level"], g["HI"], g["HeI"], g["HeII"], g["rho"], g["abun2"], g["box_size"])
    od = og.diffuse(bg["uvb"], bg["beta"])
    src, wt = np.array([N // 2, 3], dtype=np.int32), np.array([1, 2], dtype=np.int32)
    op = og.point(sp, src, wt)
    with open(os.path.join(tmp, "ftte_reference_dump.bin"), "wb") as f:
        np.array([20260001, 4, N, 2, sp["wavelength"].size, 5, 1, 0, 6, 3], dtype="<i4").tofile(f)
        np.array([g["box_size"], sp["coef_spectrum"], 3.0, *bg["uvb"], *bg["alpha"]], dtype="<f8").tofile(f)
        bg["table"][:, [0, 1, 2]].astype("<f8").tofile(f)
        bg["table"][:, [8, 9, 10]].astype("<f8").tofile(f)
        sp["wavelength"].astype("<f8").tofile(f); sp["metallicity"].astype("<f8").tofile(f)
        sp["lum"].astype("<f8").tofile(f); sp["a_dust"].astype("<f8").tofile(f)
        wt.astype("<i4").tofile(f)
        g["level"].astype("<i4").tofile(f)
        for k in ("HI", "HeI", "HeII", "rho", "abun2"):
            g[k].astype("<f8").tofile(f)
        np.full(N, 1e4).tofile(f)
        od["J"].astype("<f8").tofile(f); op["rates"].astype("<f8").tofile(f)
        for k in ("HI", "HeI", "HeII"):
            g[k].astype("<f8").tofile(f)
    with open(os.path.join(tmp, "ftte_reference_sources.bin"), "wb") as f:
        for i in range(2):
            np.array([i + 1, src[i], wt[i], op["highest_pixel_level"][i]], dtype="<i4").tofile(f)
            np.concatenate([op["ndot_remaining"][i], op["ndot_boundary"][i], op["ndot_dust"][i:i + 1]]).astype("<f8").tofile(f)
            op["ndot_spectrum"][i].astype("<f8").tofile(f)
    return compare(tmp)


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--selftest":
        sys.exit(selftest(sys.argv[2]))
    if len(sys.argv) != 2:
        raise SystemExit(__doc__)
    sys.exit(compare(sys.argv[1]))
