"""TEST INFRASTRUCTURE ONLY -- ctypes loader for the CPU oracle (oracle/libftte_oracle.so).

The oracle is a line-by-line C++ restatement of the reference hot path (see oracle/ftte_common.h for the
"parity unpinned" note).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_d = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i8 = np.ctypeslib.ndpointer(dtype=np.int8, flags="C_CONTIGUOUS")


def build(force=False):
    so = os.path.join(_HERE, "libftte_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.ftte_grid_create.restype = C.c_void_p
        L.ftte_grid_create.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_double, C.POINTER(C.c_int)]
        L.ftte_grid_destroy.argtypes = [C.c_void_p]
        L.ftte_grid_set_species.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ftte_diffuse.restype = C.c_int
        L.ftte_diffuse.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int64,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ftte_diffuse_mt.restype = C.c_int
        L.ftte_diffuse_mt.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                      C.c_void_p, C.POINTER(C.c_int64)]
        L.ftte_point.restype = C.c_int
        L.ftte_point.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p,
                                 C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.c_int64,
                                 C.POINTER(C.c_int64)]
        L.ftte_point_highest_pixel_level.restype = C.c_int
        L.ftte_point_highest_pixel_level.argtypes = [C.c_void_p, C.c_int]
        L.ftte_point_tables.restype = C.c_int
        L.ftte_point_tables.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_int,
                                        C.c_double, C.c_void_p, C.POINTER(C.c_double), C.c_void_p]
        L.ftte_direction.restype = C.c_int
        L.ftte_direction.argtypes = [C.c_int, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_double),
                                     C.POINTER(C.c_double)]
        L.ftte_pix2ang_nest.restype = C.c_int
        L.ftte_pix2ang_nest.argtypes = [C.c_int, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.ftte_rotate_indices.argtypes = [C.c_int] * 7 + [C.c_void_p]
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class OracleGrid:
    """Octree rebuilt from the flattened leaf arrays (leaf pre-order, equiSources.f90:4044-4079)."""

    def __init__(self, nx, level, HI, HeI=None, HeII=None, rho=None, abun2=None, box_size=1.0):
        self.L = lib()
        self.nx = int(nx)
        self.level = np.ascontiguousarray(level, dtype=np.int8)
        self.nleaf = int(self.level.size)
        arrs = [_f64(HI), _f64(HeI), _f64(HeII), _f64(rho), _f64(abun2)]
        st = C.c_int(0)
        self.h = self.L.ftte_grid_create(self.nx, self.nleaf, _p(self.level), *[_p(a) for a in arrs],
                                         float(box_size), C.byref(st))
        if not self.h:
            raise RuntimeError(f"ftte_grid_create failed: status {st.value}")

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ftte_grid_destroy(self.h)
            self.h = None

    def set_species(self, HI=None, HeI=None, HeII=None):
        a = [_f64(HI), _f64(HeI), _f64(HeII)]
        self.L.ftte_grid_set_species(self.h, *[_p(x) for x in a])

    def diffuse(self, uvb, beta, n_angular_level=3, ray_begin=0, ray_end=-1, trace_ray=-1):
        """Returns dict(J=[3,nleaf], nseg, status[, nb, pattern, izone, angles])."""
        uvb = _f64(uvb)
        beta = _f64(np.asarray(beta).reshape(9))
        J = np.zeros((3, self.nleaf))
        nseg = C.c_int64(0)
        nb = pat = iz = ang = None
        if trace_ray >= 0:
            nb = np.full((3, self.nleaf), -9, dtype=np.int32)
            pat = np.zeros((self.nx, 12))
            iz = np.zeros(1, dtype=np.int32)
            ang = np.zeros(2)
        st = self.L.ftte_diffuse(self.h, int(n_angular_level), _p(uvb), _p(beta), int(ray_begin), int(ray_end),
                                 _p(J[0]), _p(J[1]), _p(J[2]), C.byref(nseg), int(trace_ray),
                                 _p(nb), _p(pat), _p(iz), _p(ang))
        out = dict(J=J, nseg=nseg.value, status=st)
        if trace_ray >= 0:
            out.update(nb=nb, pattern=pat, izone=int(iz[0]), angles=ang)
        return out


def _diffuse_mt(self, uvb, beta, rays, n_angular_level=3, nthreads=1):
    """all-host-threads variant over an explicit ray list (each thread sweeps a private copy of the octree)"""
    uvb = _f64(uvb)
    beta = _f64(np.asarray(beta).reshape(9))
    rays = np.ascontiguousarray(rays, dtype=np.int32)
    J = np.zeros((3, self.nleaf))
    nseg = C.c_int64(0)
    st = self.L.ftte_diffuse_mt(self.h, int(n_angular_level), _p(uvb), _p(beta), _p(rays), int(rays.size),
                                int(nthreads), _p(J), C.byref(nseg))
    return dict(J=J, nseg=nseg.value, status=st)


OracleGrid.diffuse_mt = _diffuse_mt


def _point(self, spectra, src_leaf, src_weight, dust_approximation=0, max_pixel_level=6, rates=None, trace_cap=0):
    """Point-source pass (equiSources.f90:1256-1370).  spectra = dict(wavelength, lum[5,2,nw], metallicity[5],
    coef_spectrum, a_dust[7,5]).  Returns dict(rates[6,nleaf], ndot_remaining[nsrc,7], ndot_boundary, ndot_dust,
    ndot_spectrum[nsrc,300], nseg, status[, trace])."""
    wl = _f64(spectra["wavelength"]); lum = _f64(spectra["lum"]); met = _f64(spectra["metallicity"])
    ad = _f64(spectra["a_dust"])
    leaf = np.ascontiguousarray(src_leaf, dtype=np.int32); wt = np.ascontiguousarray(src_weight, dtype=np.int32)
    ns = int(leaf.size)
    R = np.zeros((6, self.nleaf)) if rates is None else np.ascontiguousarray(rates, dtype=np.float64).copy()
    rem = np.zeros((ns, 7)); bnd = np.zeros((ns, 7)); dust = np.zeros(ns); spec = np.zeros((ns, 300))
    nseg = C.c_int64(0); tl = C.c_int64(0)
    tr = np.zeros(max(trace_cap, 1), dtype=np.int64)
    st = self.L.ftte_point(self.h, int(wl.size), _p(wl), _p(lum), _p(met), float(spectra["coef_spectrum"]), _p(ad),
                           int(dust_approximation), int(max_pixel_level), ns, _p(leaf), _p(wt), _p(R), _p(rem), _p(bnd),
                           _p(dust), _p(spec), C.byref(nseg), _p(tr) if trace_cap else None, int(trace_cap), C.byref(tl))
    hpl = np.zeros(ns, dtype=np.int32)
    self.L.ftte_point_highest_pixel_level(_p(hpl), ns)     # same thread: the call above filled it
    out = dict(rates=R, ndot_remaining=rem, ndot_boundary=bnd, ndot_dust=dust, ndot_spectrum=spec, nseg=nseg.value,
               status=st, highest_pixel_level=hpl)
    if trace_cap:
        out["trace"] = tr[:tl.value].copy()
    return out


OracleGrid.point = _point


def chemistry(nx, box_size, level, rho, tgas, HI, HeI, HeII, ktab, rates=None, J=None, ksi=None, uniform=None):
    """solveRateEquations (equiSources.f90:3459-3677) on leaf arrays.  ktab = dict(k=[6, nratec], logtem0, logtem9,
    dlogtem).  Returns dict(HI, HeI, HeII, max_change, status)."""
    L = lib()
    level = np.ascontiguousarray(level, dtype=np.int8)
    n = int(level.size)
    out = [np.ascontiguousarray(a, dtype=np.float64).copy() for a in (HI, HeI, HeII)]
    k = _f64(ktab["k"])
    ksi_a = _f64(np.zeros(6) if ksi is None else ksi)
    uni = _f64(np.zeros(4) if uniform is None else uniform)
    mc = C.c_double(0)
    L.ftte_chemistry.restype = C.c_int
    L.ftte_chemistry.argtypes = [C.c_int64, C.c_int, C.c_double] + [C.c_void_p] * 10 + [C.c_int, C.c_double, C.c_double,
                                                                                       C.c_double, C.c_void_p,
                                                                                       C.POINTER(C.c_double)]
    st = L.ftte_chemistry(n, int(nx), float(box_size), _p(level), _p(_f64(rho)), _p(_f64(tgas)), _p(out[0]), _p(out[1]),
                          _p(out[2]), _p(_f64(rates)), _p(_f64(J)), _p(ksi_a), _p(uni), int(k.shape[1]),
                          float(ktab["logtem0"]), float(ktab["logtem9"]), float(ktab["dlogtem"]), _p(k), C.byref(mc))
    return dict(HI=out[0], HeI=out[1], HeII=out[2], max_change=mc.value, status=st)


def compute_mass(nx, box_size, level, HI, rho):
    """computeMass (equiSources.f90:4369-4393): (neutralHydrogenMass, totalHydrogenMass) in solar masses, summed
    serially in leaf order"""
    L = lib()
    level = np.ascontiguousarray(level, dtype=np.int8)
    out = np.zeros(2)
    L.ftte_compute_mass.restype = None
    L.ftte_compute_mass.argtypes = [C.c_int64, C.c_int, C.c_double] + [C.c_void_p] * 4
    L.ftte_compute_mass(int(level.size), int(nx), float(box_size), _p(level), _p(_f64(HI)), _p(_f64(rho)), _p(out))
    return float(out[0]), float(out[1])


def set_portable_math(on):
    """point path: evaluate exp/log with radiativetransfer_b200/csrc/portable_math.h (IEEE +,*,/,fma only) instead
    of libm -- the same source the CUDA kernels use in FAITHFUL mode, so deposits can be compared bit for bit"""
    lib().ftte_set_portable_math(int(bool(on)))


def pm_eval(x):
    """(pm_exp(x), pm_log(x)) of csrc/portable_math.h compiled for the host"""
    L = lib()
    L.ftte_pm_eval.restype = None
    L.ftte_pm_eval.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    x = _f64(x)
    e, l = np.empty_like(x), np.empty_like(x)
    L.ftte_pm_eval(int(x.size), _p(x), _p(e), _p(l))
    return e, l


def point_tables(spectra, i_metal, coef_metal):
    wl = _f64(spectra["wavelength"]); lum = _f64(spectra["lum"]); met = _f64(spectra["metallicity"])
    ad = _f64(spectra["a_dust"])
    out = np.zeros((6, 11 ** 4)); tot = C.c_double(0); sig = np.zeros((5, 300))
    st = lib().ftte_point_tables(int(wl.size), _p(wl), _p(lum), _p(met), float(spectra["coef_spectrum"]), _p(ad),
                                 int(i_metal), float(coef_metal), _p(out), C.byref(tot), _p(sig))
    return dict(tables=out, total_integral=tot.value, output_sigma=sig, status=st)


def uvb_tables(redshift=3.0, uvb_coefficient=1.0, nfreq=400, freqdel=float(np.float32(0.02))):
    """equiSources.f90:198-246 + powerSpectrumIndex (:4985-5043) + uvbBetaTable.f90: dict(uvb[3], alpha[3], extra[8] =
    uvbStellar1..3, uvbQuasar1..3, uniformQuasar, uniformStellar; table[3, 19] = beta24..31, ksi24..31, gammaHI, gammaHeI,
    gammaHeII per group; status)"""
    L = lib()
    L.ftte_uvb.restype = C.c_int
    L.ftte_uvb.argtypes = [C.c_double, C.c_double, C.c_int, C.c_double] + [C.c_void_p] * 4
    uvb, alpha, extra, tab = np.zeros(3), np.zeros(3), np.zeros(8), np.zeros((3, 19))
    st = L.ftte_uvb(float(redshift), float(uvb_coefficient), int(nfreq), float(freqdel), _p(uvb), _p(alpha), _p(extra),
                    _p(tab))
    return dict(uvb=uvb, alpha=alpha, extra=extra, table=tab, status=st)


def power_spectrum_index(uvb1, alpha1, uvb2, alpha2, nug, nugplus, bound):
    L = lib()
    L.ftte_power_spectrum_index.restype = C.c_int
    L.ftte_power_spectrum_index.argtypes = [C.c_double] * 6 + [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    tot, al = C.c_double(0), C.c_double(0)
    st = L.ftte_power_spectrum_index(uvb1, alpha1, uvb2, alpha2, nug, nugplus, int(bool(bound)), C.byref(tot), C.byref(al))
    return st, tot.value, al.value


def direction(n_angular_level, iray):
    iz, phi, th = C.c_int32(0), C.c_double(0), C.c_double(0)
    st = lib().ftte_direction(int(n_angular_level), int(iray), C.byref(iz), C.byref(phi), C.byref(th))
    return st, iz.value, phi.value, th.value


def pix2ang_nest(nside, ipix):
    phi, th = C.c_double(0), C.c_double(0)
    st = lib().ftte_pix2ang_nest(int(nside), int(ipix), C.byref(phi), C.byref(th))
    return st, phi.value, th.value


def rotate_indices(i, j, k, nx, ny, nz, izone):
    out = np.zeros(3, dtype=np.int32)
    lib().ftte_rotate_indices(i, j, k, nx, ny, nz, izone, _p(out))
    return tuple(int(v) for v in out)
