"""Row a10 of SURVEY.md section 8: uvbBetaTable (uvbBetaTable.f90:31-296), powerSpectrumIndex (equiSources.f90:4985-5043)
and the band amplitudes (equiSources.f90:198-246).

Three evaluations are compared: the oracle's line-by-line restatement (oracle/ftte_uvb.cpp), the product's host code
(csrc/uvb_host.cpp, behind rtb200_uvb_*), and an independent numpy evaluation written here from the formulas (vectorised
sums, so it differs from the serial sums by rounding only).  CPU only: these tables are host work."""
import ctypes as C

import numpy as np
import pytest

from radiativetransfer_b200 import workloads as W

f32 = lambda x: float(np.float32(x))
NU = [f32(13.598), f32(24.587), f32(54.418)]
PI = f32(3.141592654)
EV_TO_ERG = 1.60217646e-12
EV_TO_HZ = EV_TO_ERG / f32(6.6260693e-27)


def numpy_sigmas(nu):
    """the 8 cross-sections of uvbBetaTable.f90:35-101 on an energy grid [eV], single-precision literals widened"""
    s = np.zeros((8, nu.size))
    for row, (s0, e0) in ((0, (f32(6.3e-18), NU[0])), (1, (f32(1.58e-18), NU[2]))):
        m = nu > e0
        d = np.sqrt(nu[m] / e0 - 1)
        s[row, m] = s0 * (e0 / nu[m]) ** 4 * np.exp(4.0 - 4.0 * np.arctan(d) / d) / (1 - np.exp(-2.0 * PI / d))
    m = nu > NU[1]
    r = nu[m] / NU[1]
    s[2, m] = f32(7.42e-18) * (f32(1.66) * r ** f32(-2.05) - f32(0.66) * r ** f32(-3.05))
    m = nu > f32(0.755)
    s[3, m] = f32(2.11e-16) * (nu[m] - f32(0.755)) ** 1.5 / nu[m] ** 3
    m = (nu > f32(2.65)) & (nu <= f32(11.27))
    x = nu[m]
    s[4, m] = 10.0 ** (f32(-40.97) + f32(6.03) * x - f32(0.504) * x ** 2 + f32(1.387e-2) * x ** 3)
    m = (nu > f32(11.27)) & (nu < f32(21.0))
    x = nu[m]
    s[4, m] = 10.0 ** (f32(-30.26) + f32(2.79) * x - f32(0.184) * x ** 2 + f32(3.535e-3) * x ** 3)
    m = (nu > f32(15.42)) & (nu <= f32(16.5))
    s[5, m] = f32(6.2e-18) * nu[m] - f32(9.4e-17)
    m = (nu > f32(16.5)) & (nu <= f32(17.7))
    s[5, m] = f32(1.4e-18) * nu[m] - f32(1.48e-17)
    m = nu > f32(17.7)
    s[5, m] = f32(2.5e-14) * nu[m] ** f32(-2.71)
    m = (nu >= f32(30.0)) & (nu < f32(70.0))
    x = nu[m]
    s[6, m] = 10.0 ** (f32(-16.926) - f32(4.528e-2) * x + f32(2.238e-4) * x ** 2 + f32(4.245e-7) * x ** 3)
    m = (nu > f32(11.27)) & (nu < NU[0])
    s[7, m] = f32(3.71e-18)
    return s


def numpy_beta_table(alpha, nfreq=400, freqdel=f32(0.02)):
    nu = 10.0 ** (np.arange(nfreq) * freqdel)
    sig = numpy_sigmas(nu)
    dnu = np.diff(nu, prepend=nu[0])
    out = np.zeros((3, 19))
    inband = [(nu >= NU[0]) & (nu <= NU[1]), (nu >= NU[1]) & (nu <= NU[2]), nu >= NU[2]]
    shape = [(1.0 - (NU[1] / NU[0]) ** (1.0 - alpha[0])) / (alpha[0] - 1.0),
             (1.0 - (NU[2] / NU[1]) ** (1.0 - alpha[1])) / (alpha[1] - 1.0), 1.0 / (alpha[2] - 1.0)]
    for g in range(3):
        m = inband[g].copy()
        m[0] = False                                  # the reference's loop starts at i = 2
        dt = (nu[m] / NU[g]) ** (-alpha[g]) * dnu[m]
        dte = dt * EV_TO_HZ / (nu[m] * EV_TO_ERG)
        out[g, :8] = (sig[:, m] * dt).sum(axis=1) / (shape[g] * NU[g])
        out[g, 8:16] = (sig[:, m] * dte).sum(axis=1)
        out[g, 16] = np.sum(dte * (nu[m] - NU[0]) * EV_TO_ERG * sig[0, m])
        if g >= 1:
            out[g, 17] = np.sum(dte * (nu[m] - NU[1]) * EV_TO_ERG * sig[2, m])
        if g == 2:
            out[g, 18] = np.sum(dte * (nu[m] - NU[2]) * EV_TO_ERG * sig[1, m])
    return out


def numpy_amplitudes(z, coef=1.0):
    """equiSources.f90:198-242: band amplitudes of the stellar and quasar components"""
    damp = 1.0 + (7.0 / (1.0 + z)) ** 4
    stellar99 = 1.0 / damp * np.exp(-((z / 4.0) ** 3))
    pascal02 = f32(0.0188) * np.exp(-((z - 0.5) ** 2) / (1.0 + f32(0.0625) * (z + f32(2.09)) ** f32(2.075))) * \
        (1.0 + z) ** f32(3.35)
    step = 0.5 * (np.tanh((z - f32(4.2)) * 1.5) + 1.0)
    stellar02 = (1.0 - step) * stellar99 + step * pascal02
    quasar02 = 10.0 / damp * np.exp(-((z / 2.5) ** 3))
    gaussian = np.exp(-(((z - 4.5) / 2.0) ** 2)) * f32(0.3)
    newQ = gaussian * stellar02 + (1.0 - gaussian) * quasar02
    newS = (1.0 - gaussian) * stellar02 + gaussian * quasar02
    newS = (1.0 - 0.5 * (np.tanh((z - 14.0) * 0.5) + 1.0)) * newS
    aQ, aS = f32(1.8), 5.0
    s = [newS * 1e-21 * coef]
    q = [newQ * 1e-21 * coef]
    for g in (1, 2):
        s.append(s[-1] * (NU[g] / NU[g - 1]) ** (-aS))
        q.append(q[-1] * (NU[g] / NU[g - 1]) ** (-aQ))
    return np.array(s), np.array(q), aS, aQ


def band_integral(u, a, g):
    """integral of u (nu/nu_g)^-a d(nu/nu_g) over band g (the quantity powerSpectrumIndex matches)"""
    if g < 2:
        return u / (a - 1.0) * (1.0 - (NU[g] / NU[g + 1]) ** (a - 1.0))
    return u / (a - 1.0)


@pytest.mark.parametrize("z", [0.0, 3.0, 6.5])
def test_oracle_amplitudes_and_slopes(oracle, z):
    o = oracle.uvb_tables(z, 0.7)
    assert o["status"] == 0
    s, q, aS, aQ = numpy_amplitudes(z, 0.7)
    assert np.allclose(o["extra"][:3], s, rtol=1e-13, atol=0) and np.allclose(o["extra"][3:6], q, rtol=1e-13, atol=0)
    assert np.allclose(o["uvb"], s + q, rtol=1e-14, atol=0)
    for g in range(3):
        # defining property of powerSpectrumIndex: the single power law reproduces the two-component band integral
        lhs = band_integral(o["uvb"][g], o["alpha"][g], g)
        rhs = band_integral(s[g], aS, g) + band_integral(q[g], aQ, g)
        assert abs(lhs / rhs - 1.0) < 1e-7                     # the reference iterates to |d alpha| < 1e-8
        assert min(aS, aQ) < o["alpha"][g] < max(aS, aQ)


def test_power_spectrum_index_wrong_sign_is_reported(oracle):
    # equal slopes make both bracket ends coincide: 'wrong sign ... stop' (equiSources.f90:5016-5019)
    st, _, _ = oracle.power_spectrum_index(1e-21, 2.0, 1e-21, 2.0, NU[0], NU[1], True)
    assert st != 0
    st, tot, al = oracle.power_spectrum_index(1e-21, 5.0, 3e-21, f32(1.8), NU[0], NU[1], True)
    assert st == 0 and tot == 4e-21 and f32(1.8) < al < 5.0


@pytest.mark.parametrize("z", [0.0, 3.0, 6.5])
def test_oracle_beta_table_against_numpy(oracle, z):
    o = oracle.uvb_tables(z)
    ref = numpy_beta_table(o["alpha"])
    m = ref != 0
    assert np.array_equal(o["table"] != 0, m)
    assert np.max(np.abs(o["table"][m] / ref[m] - 1.0)) < 1e-12
    # structure: HeII is not ionised below nu3, HeI not below nu2; group cross-sections are positive and below the threshold value
    t = o["table"]
    assert t[0, 1] == 0 and t[1, 1] == 0 and t[0, 2] == 0 and t[2, 1] > 0
    assert 0 < t[0, 0] < f32(6.3e-18) and 0 < t[2, 1] < f32(1.58e-18)
    assert np.all(t[:, 16] > 0) and t[0, 17] == 0 and t[0, 18] == 0 and t[1, 18] == 0


@pytest.mark.parametrize("z,coef", [(0.0, 1.0), (3.0, 1.0), (3.0, 0.25), (6.5, 2.0), (9.0, 1.0)])
def test_product_tables_equal_oracle(build_product, oracle, z, coef):
    o = oracle.uvb_tables(z, coef)
    p = W.uvb_background(z, coef)
    # same formulas, same libm, separately written: identical bits expected
    assert np.array_equal(p["uvb"], o["uvb"]) and np.array_equal(p["alpha"], o["alpha"])
    assert np.array_equal(p["table"], o["table"])
    t = o["table"]
    assert np.array_equal(p["beta"], t[:, [0, 2, 1]])                 # [group][beta24, beta26, beta25]
    assert np.array_equal(p["ksi24"], t[:, 8]) and p["ksi25"][0] == t[2, 9]
    assert np.array_equal(p["ksi26"], t[[1, 2], 10])


def test_product_entry_points_argument_checks(build_product):
    from radiativetransfer_b200 import _lib
    L = _lib.lib()
    a = np.zeros(3)
    assert L.rtb200_uvb_amplitudes(3.0, 1.0, None, a.ctypes.data_as(C.c_void_p)) == 12
    assert L.rtb200_uvb_beta_table(1, 0.02, a.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p)) == 12
    # a bin grid other than the reference's is legal (coarser grid, same structure)
    t = np.zeros(57)
    al = np.array([2.5, 2.2, 1.9])
    assert L.rtb200_uvb_beta_table(200, 0.04, al.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p)) == 0
    ref = numpy_beta_table(al, 200, 0.04).ravel()
    m = ref != 0
    assert np.max(np.abs(t[m] / ref[m] - 1.0)) < 1e-12
