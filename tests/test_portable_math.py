"""csrc/portable_math.h is shared by the CUDA kernels (FAITHFUL point-source mode) and by the oracle's `portable`
switch, so a comparison between those two cannot see an error in the header itself.  Here the header is pinned by an
independent evaluation: Python's `decimal` at 60 significant digits (correctly rounded exp / ln), measured in ulps."""
import ctypes as C
from decimal import Decimal, getcontext

import numpy as np
import pytest


def _ulp_errors(x, got, fn):
    getcontext().prec = 60
    errs = np.zeros(x.size)
    for i, (xi, gi) in enumerate(zip(x.tolist(), got.tolist())):
        exact = fn(Decimal(xi))
        ulp = Decimal(float(np.spacing(abs(gi)))) if gi != 0 else Decimal(5e-324)
        errs[i] = float(abs(Decimal(gi) - exact) / ulp)
    return errs


def _sample(seed=11):
    rng = np.random.default_rng(seed)
    xe = np.concatenate([rng.uniform(-700.0, 700.0, 4000), rng.uniform(-40.0, 5.0, 6000), -10.0 ** rng.uniform(-12, 2, 3000),
                         rng.uniform(-1e-3, 1e-3, 1000), np.array([0.0, 1.0, -1.0, 0.5 * np.log(2.0), -0.5 * np.log(2.0)])])
    xl = np.concatenate([10.0 ** rng.uniform(-300, 300, 4000), 10.0 ** rng.uniform(-30, 45, 6000),
                         rng.uniform(0.5, 2.0, 3000), 1.0 + rng.uniform(-1e-6, 1e-6, 1000),
                         np.array([1.0, 2.0, np.sqrt(2.0), np.nextafter(np.sqrt(2.0), 0), 5e-324, 2.2250738585072014e-308])])
    return xe, xl


def test_host_build_against_60_digit_evaluation(oracle):
    xe, xl = _sample()
    e, _ = oracle.pm_eval(xe)
    _, l = oracle.pm_eval(xl)
    ee = _ulp_errors(xe, e, lambda d: d.exp())
    el = _ulp_errors(xl, l, lambda d: d.ln())
    print(f"pm_exp: max {ee.max():.3f} ulp, mean {ee.mean():.3f}; pm_log: max {el.max():.3f} ulp, mean {el.mean():.3f}")
    assert ee.max() < 1.0          # the accuracy the header states: < 1 ulp
    assert el.max() < 1.5          # < 1.5 ulp
    # special values
    sp_e, _ = oracle.pm_eval(np.array([-746.0, 710.0, np.nan, 0.0]))
    assert sp_e[0] == 0.0 and np.isinf(sp_e[1]) and np.isnan(sp_e[2]) and sp_e[3] == 1.0
    _, sp_l = oracle.pm_eval(np.array([0.0, -1.0, np.inf, 1.0]))
    assert sp_l[0] == -np.inf and np.isnan(sp_l[1]) and sp_l[2] == np.inf and sp_l[3] == 0.0


def test_against_libm_on_the_ranges_the_point_path_uses(oracle):
    # table sums: exp(-tau), tau in [0, 150]; lookups: log of sums in [1e-300, 1e60]
    rng = np.random.default_rng(5)
    x = -rng.uniform(0.0, 150.0, 200000)
    e, _ = oracle.pm_eval(x)
    assert np.max(np.abs(e / np.exp(x) - 1.0)) < 4.5e-16
    y = 10.0 ** rng.uniform(-300, 60, 200000)
    _, l = oracle.pm_eval(y)
    assert np.max(np.abs(l - np.log(y)) / np.maximum(np.abs(np.log(y)), 1e-300)) < 4.5e-16


@pytest.mark.gpu
def test_device_build_is_bit_identical_to_host_build(build_product, oracle):
    import radiativetransfer_b200 as rt
    xe, xl = _sample(seed=12)
    x = np.concatenate([xe, xl, np.array([-746.0, 710.0, 0.0, -0.0, np.inf])])
    t = rt.Transport(device=0)
    e, l = np.empty_like(x), np.empty_like(x)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    st = t.L.rtb200_debug_portable_math(t.h, int(x.size), p(x), p(e), p(l))
    assert st == 0
    he, hl = oracle.pm_eval(x)
    assert np.array_equal(e.view(np.int64), he.view(np.int64))
    assert np.array_equal(l.view(np.int64), hl.view(np.int64))
    t.close()
