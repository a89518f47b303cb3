"""GPU parity tests of the point-source path: CUDA ray casting / deposition through the C-ABI vs the CPU oracle on the
same seeded inputs.  Bars (BASELINE.json north_star): bit-exact ray-to-cell traversal order; <= 1e-9 relative L-infinity
on the per-cell rates (fp64).

The deposits are differences R(d) - R(d+tau) of exponentials of interpolated logarithms (equiSources.f90:3247-3260):
for a short segment they cancel, and a last-bit difference between two libm implementations is amplified by ~1/tau.
The parity proper therefore runs with the SAME exp/log source on both sides (csrc/portable_math.h: IEEE +,*,/,fma
only; device: RTB200_MATH_FAITHFUL, oracle: set_portable_math(1)), where every deposit is bit-identical and only the
order of the per-cell additions differs.  Against the libm oracle (what gfortran links) and for RTB200_MATH_FAST the
same tolerance is applied plus the documented conditioning floor of the reference formula itself."""
import numpy as np
import pytest

from conftest import rel_err
from radiativetransfer_b200 import workloads as W

pytestmark = pytest.mark.gpu
TOL = 1e-9
S24 = float(np.float32(6.3e-18))


@pytest.fixture(scope="module")
def rt(build_product):
    import radiativetransfer_b200 as rt
    return rt


@pytest.fixture()
def engine(rt):
    t = rt.Transport(device=0, math=rt.MATH_FAITHFUL)
    yield t
    t.close()


@pytest.fixture(scope="module")
def spectra():
    return W.synthetic_spectra()


@pytest.fixture()
def portable(oracle):
    oracle.set_portable_math(True)
    yield oracle
    oracle.set_portable_math(False)


def _set(t, g):
    t.set_grid(g["nx"], g["level"], g["HI"], g["HeI"], g["HeII"], g["rho"], g["abun2"], g["box_size"])


def _ograd(oracle, g):
    return oracle.OracleGrid(g["nx"], g["level"], g["HI"], g["HeI"], g["HeII"], g["rho"], g["abun2"], g["box_size"])


def _cell_err(a, b):
    """relative L-infinity per cell, over cells the oracle deposits into"""
    m = b != 0
    assert np.all(a[~m] == 0)
    return float(np.max(np.abs(a[m] - b[m]) / np.abs(b[m]))) if m.any() else 0.0


def _by_ray(trace):
    """stable sort of an oracle trace (depth-first order) by (pixel level, pixel): per-ray segment order is kept"""
    key = ((trace >> 28) & 0xF) << 24 | ((trace >> 8) & 0xFFFFF)
    return trace[np.argsort(key, kind="stable")]


def test_tables_match_oracle(engine, portable, spectra):
    g = W.uniform_grid(4, seed=1)
    _set(engine, g)
    for im, cm in [(1, 0.0), (2, 0.25), (4, 1.0)]:
        T = engine.point_tables(spectra, im, cm)
        o = portable.point_tables(spectra, im, cm)
        assert o["status"] == 0
        assert np.array_equal(T, o["tables"])          # same sums, same order, same exp: bit-identical
    portable.set_portable_math(False)
    o = portable.point_tables(spectra, 2, 0.25)        # libm oracle: 399-term sums of 1-ulp-different exponentials
    assert rel_err(engine.point_tables(spectra, 2, 0.25), o["tables"]) < 1e-13


@pytest.mark.parametrize("n,src,maxlev,dust,helium", [
    (8, (3, 4, 2), 1, 0, True), (12, (6, 6, 6), 3, 0, True), (16, (8, 7, 9), 6, 0, False), (16, (2, 13, 8), 6, 1, True),
    (10, (0, 0, 0), 5, 2, True), (32, (16, 16, 16), 6, 0, False)])
def test_uniform_parity_portable(engine, portable, spectra, n, src, maxlev, dust, helium):
    g = W.uniform_grid(n, seed=n + maxlev, tau_lo=1e-3, tau_hi=1.0, beta24=S24, helium=helium)
    g["abun2"] = np.random.default_rng(n).uniform(1e-3, 4e-2, n ** 3)
    _set(engine, g)
    leaf = (src[0] * n + src[1]) * n + src[2]
    r = engine.point(spectra, [leaf], [2], dust_approximation=dust, max_pixel_level=maxlev)
    o = _ograd(portable, g).point(spectra, [leaf], [2], dust_approximation=dust, max_pixel_level=maxlev)
    assert o["status"] == 0
    assert r["nseg"] == o["nseg"]
    for i in range(6):
        assert _cell_err(r["rates"][i], o["rates"][i]) < 1e-12, i   # identical deposits, different summation order
    assert rel_err(r["ndot_remaining"], o["ndot_remaining"], floor=1e-300) < 1e-12
    assert rel_err(r["ndot_boundary"], o["ndot_boundary"], floor=1e-300) < 1e-12


@pytest.mark.parametrize("n,src,maxlev", [(8, (3, 4, 2), 1), (12, (6, 5, 6), 4), (16, (8, 7, 9), 6)])
def test_uniform_traversal_bit_exact(engine, portable, spectra, n, src, maxlev):
    g = W.uniform_grid(n, seed=3, tau_lo=1e-3, tau_hi=0.5, beta24=S24)
    _set(engine, g)
    leaf = (src[0] * n + src[1]) * n + src[2]
    r = engine.point(spectra, [leaf], [1], max_pixel_level=maxlev, trace_cap=400000)
    o = _ograd(portable, g).point(spectra, [leaf], [1], max_pixel_level=maxlev, trace_cap=400000)
    assert o["status"] == 0 and r["nseg"] == o["nseg"] == o["trace"].size
    assert np.array_equal(r["trace"], _by_ray(o["trace"]))   # (leaf, pixel level, pixel, exit face) of every segment


def test_amr_parity_and_traversal(engine, portable, spectra):
    n = 8
    g = W.nested_grid(n, 3, W.central_box_refine(0.2, 0.8, levels=3), seed=4, tau_lo=1e-3, tau_hi=0.5, beta24=S24)
    _set(engine, g)
    cx, cy, cz = g["centres"]
    srcs = [int(np.argmin((cx - a) ** 2 + (cy - b) ** 2 + (cz - c) ** 2))
            for a, b, c in [(0.51, 0.52, 0.47), (0.1, 0.9, 0.3), (0.79, 0.5, 0.21)]]
    og = _ograd(portable, g)
    for dust in (0, 2):
        r = engine.point(spectra, srcs, [1, 3, 2], dust_approximation=dust)
        o = og.point(spectra, srcs, [1, 3, 2], dust_approximation=dust)
        assert o["status"] == 0 and r["nseg"] == o["nseg"]
        for i in range(6):
            assert _cell_err(r["rates"][i], o["rates"][i]) < 1e-12
        assert rel_err(r["ndot_remaining"], o["ndot_remaining"], floor=1e-300) < 1e-12
        assert rel_err(r["ndot_boundary"], o["ndot_boundary"], floor=1e-300) < 1e-12
    r = engine.point(spectra, srcs[:1], [1], trace_cap=3000000)
    o = og.point(spectra, srcs[:1], [1], trace_cap=3000000)
    assert np.array_equal(r["trace"], _by_ray(o["trace"]))


def test_against_libm_oracle_and_fast_mode(rt, engine, oracle, spectra):
    """the reference's own libm (glibc) and the FAST reformulation: same result up to the conditioning of
    R(d) - R(d+tau); cells are compared at 1e-9 relative plus 2e-13 of the undepleted rate a ray carries into the cell
    (= the rounding noise of exp(interpolated log) that the subtraction exposes)"""
    n = 16
    g = W.uniform_grid(n, seed=9, tau_lo=1e-3, tau_hi=1.0, beta24=S24)
    _set(engine, g)
    leaf = (8 * n + 8) * n + 8
    o = _ograd(oracle, g).point(spectra, [leaf], [1])
    assert o["status"] == 0
    T = engine.point_tables(spectra, *_bracket(spectra, g["abun2"][leaf])).reshape(6, -1)
    r_faith = engine.point(spectra, [leaf], [1])
    engine.set_math(rt.MATH_FAST)
    r_fast = engine.point(spectra, [leaf], [1])
    for r in (r_faith, r_fast):
        assert r["nseg"] == o["nseg"]
        for i, t in zip(range(6), (0, 2, 1, 3, 5, 4)):
            floor = 2e-13 * T[t, 0]     # R_r(0): the largest value either side of the subtraction can take
            d = np.abs(r["rates"][i] - o["rates"][i])
            assert np.all(d <= TOL * np.abs(o["rates"][i]) + floor), (i, float(np.max(d / np.maximum(np.abs(o["rates"][i]), 1e-300))))
        assert np.allclose(r["rates"].sum(axis=1), o["rates"].sum(axis=1), rtol=1e-11, atol=0)
        assert rel_err(r["ndot_remaining"], o["ndot_remaining"], floor=1e-300) < 1e-11


def _bracket(spectra, abun2):
    t = np.log10(abun2) if abun2 > 1e-20 else -20.0
    met = spectra["metallicity"]
    m = 1
    while t > met[m]:
        m += 1
        if m + 1 == 5:
            break
    return m, float(np.clip((t - met[m - 1]) / (met[m] - met[m - 1]), 0, 1))


def test_sources_accumulate_and_shard(engine, portable, spectra):
    """rates accumulate over calls like the reference's cell fields; sharding the source list across ranks and adding
    the per-rank rates gives the one-call result"""
    n = 12
    g = W.uniform_grid(n, seed=12, tau_lo=1e-2, tau_hi=1.0, beta24=S24)
    _set(engine, g)
    rng = np.random.default_rng(5)
    leaves = rng.choice(n ** 3, 7, replace=False).astype(np.int32)
    wts = rng.integers(0, 4, 7).astype(np.int32)   # includes weight 0 (skipped, equiSources.f90:1264)
    full = engine.point(spectra, leaves, wts)
    a = engine.point(spectra, leaves[:3], wts[:3])
    b = engine.point(spectra, leaves[3:], wts[3:], rates=a["rates"])
    assert b["nseg"] + a["nseg"] == full["nseg"]
    assert rel_err(b["rates"], full["rates"], floor=1e-300) < 1e-12
    engine.set_tuning(point_refill=1)              # last level: lanes take pixels from a queue (other ray -> lane map)
    d = engine.point(spectra, np.repeat(leaves, 3), np.repeat(wts, 3))   # enough rays for the queue to be used
    engine.set_tuning(point_refill=0)
    e = engine.point(spectra, np.repeat(leaves, 3), np.repeat(wts, 3))
    assert d["nseg"] == e["nseg"] == 3 * full["nseg"]
    assert rel_err(d["rates"], e["rates"], floor=1e-300) < 1e-12
    engine.set_tuning(point_batch=2)               # batching of sources does not change anything
    c = engine.point(spectra, leaves, wts)
    assert rel_err(c["rates"], full["rates"], floor=1e-300) < 1e-12
    assert np.array_equal(c["ndot_boundary"], full["ndot_boundary"])
    o = _ograd(portable, g).point(spectra, leaves, wts)
    for i in range(6):
        assert _cell_err(full["rates"][i], o["rates"][i]) < 1e-12


def test_photon_conservation_256_scale_properties(engine, spectra):
    """size-independent properties at a larger size (64^3 + one refined level), no oracle: opaque box absorbs
    R(0) * weight; empty box deposits nothing and every ray ends on the boundary"""
    n = 64
    g = W.nested_grid(n, 1, W.central_box_refine(0.375, 0.625, levels=1), seed=6, tau_lo=3.0, tau_hi=8.0, beta24=S24)
    g["HeI"][:] = 0; g["HeII"][:] = 0
    _set(engine, g)
    cx, cy, cz = g["centres"]
    src = int(np.argmin((cx - 0.5) ** 2 + (cy - 0.5) ** 2 + (cz - 0.5) ** 2))
    r = engine.point(spectra, [src], [5])
    T = engine.point_tables(spectra, *_bracket(spectra, g["abun2"][src]))
    assert np.isclose(r["rates"][0].sum(), 5 * T[0, 0], rtol=1e-11)
    assert np.isclose(r["rates"][3].sum(), 5 * T[3, 0], rtol=1e-11)
    assert np.all(r["rates"][[1, 2, 4, 5]] == 0)
    engine.update_species(np.zeros_like(g["HI"]), None, None)
    r = engine.point(spectra, [src], [5])
    assert np.all(r["rates"] == 0)
    tot = r["ndot_remaining"][0] + r["ndot_boundary"][0]
    assert np.all(tot <= 5 * (1 + 1e-12)) and np.all(tot > 4.8)


def test_golden_fixtures(rt, engine, spectra):
    """committed oracle outputs (tests/golden/, libm oracle): rates within 1e-9 relative plus the conditioning floor,
    diagnostics within 1e-11, traversal bit-exact"""
    import os
    from conftest import ROOT
    f = np.load(os.path.join(ROOT, "tests", "golden", "point_uniform_10.npz"))
    g = W.uniform_grid(10, seed=int(f["seed"]), tau_lo=1e-2, tau_hi=1.0, beta24=S24)
    g["abun2"] = f["abun2"]
    _set(engine, g)
    # conditioning floor of R(d) - R(d + tau): 2e-13 of the undepleted rate the sources emit, sum_s weight_s * R_r(0)
    scale = np.zeros((6, 1))
    for leaf, w in zip(f["src"], f["wt"]):
        T = engine.point_tables(spectra, *_bracket(spectra, g["abun2"][leaf]))
        scale += w * T[[0, 2, 1, 3, 5, 4], 0][:, None]           # rates are ordered 24, 25, 26; tables reaction 1, 2, 3
    for mode in (rt.MATH_FAITHFUL, rt.MATH_FAST):
        engine.set_math(mode)
        r = engine.point(spectra, f["src"], f["wt"], dust_approximation=int(f["dust"]))
        assert r["nseg"] == int(f["nseg"])
        assert np.all(np.abs(r["rates"] - f["rates"]) <= TOL * np.abs(f["rates"]) + 2e-13 * scale)
        assert rel_err(r["ndot_remaining"], f["ndot_remaining"], floor=1e-300) < 1e-11
        assert rel_err(r["ndot_dust"], f["ndot_dust"], floor=1e-300) < 1e-11
        assert rel_err(r["ndot_boundary"], f["ndot_boundary"], floor=1e-300) < 1e-12   # same terms, other order
    f = np.load(os.path.join(ROOT, "tests", "golden", "point_amr_6.npz"))
    g = W.nested_grid(6, 2, W.central_box_refine(0.2, 0.8, levels=2), seed=int(f["seed"]), tau_lo=1e-2, tau_hi=0.5,
                      beta24=S24)
    _set(engine, g)
    engine.set_math(rt.MATH_FAITHFUL)
    r = engine.point(spectra, f["src"], f["wt"], trace_cap=400000)
    assert r["nseg"] == int(f["nseg"])
    # the oracle traces source by source, depth first; the device sorts by (source, pixel level, pixel, segment)
    tr, n0 = f["trace"], int(f["nseg_source0"])
    assert np.array_equal(r["trace"], np.concatenate([_by_ray(tr[:n0]), _by_ray(tr[n0:])]))


def test_segmented_deposition_is_deterministic_and_atomic_free(rt, engine, portable, spectra):
    """set_tuning(point_deposit=1): records + radix sort by (leaf, ray, segment) + one thread per cell.  Same deposits
    as the RED.ADD path (<= 1e-12 per cell: only the order of the additions differs), bit-identical from run to run
    and independent of the batching of sources."""
    n = 12
    g = W.nested_grid(n, 1, W.central_box_refine(0.25, 0.75, levels=1), seed=21, tau_lo=1e-2, tau_hi=1.0, beta24=S24)
    _set(engine, g)
    rng = np.random.default_rng(8)
    leaves = rng.choice(g["level"].size, 5, replace=False).astype(np.int32)
    wts = np.array([1, 2, 0, 3, 1], dtype=np.int32)
    engine.set_tuning(point_deposit=0)
    atom = engine.point(spectra, leaves, wts, dust_approximation=1)
    engine.set_tuning(point_deposit=1)
    seg1 = engine.point(spectra, leaves, wts, dust_approximation=1)
    seg2 = engine.point(spectra, leaves, wts, dust_approximation=1)
    assert seg1["nseg"] == atom["nseg"]
    assert np.array_equal(seg1["rates"], seg2["rates"])
    for i in range(6):
        assert _cell_err(seg1["rates"][i], atom["rates"][i]) < 1e-12
    o = _ograd(portable, g).point(spectra, leaves, wts, dust_approximation=1)
    for i in range(6):
        assert _cell_err(seg1["rates"][i], o["rates"][i]) < 1e-12
    # accumulation into existing rates, and a record buffer that is too small is reported, not truncated
    seg3 = engine.point(spectra, leaves, wts, dust_approximation=1, rates=seg1["rates"])
    assert rel_err(seg3["rates"], 2 * seg1["rates"], floor=1e-300) < 1e-12
    engine.set_tuning(point_record_cap=1000)
    with pytest.raises(rt.RTB200Error):
        engine.point(spectra, leaves, wts, dust_approximation=1)
    engine.set_tuning(point_record_cap=0, point_deposit=0)


def test_edge_cases(rt, engine, spectra):
    """empty source list, all weights zero, a single pixel level, a source in a corner leaf of a 1^3 grid, bad arguments"""
    g = W.uniform_grid(6, seed=2, tau_lo=1e-2, tau_hi=1.0, beta24=S24)
    _set(engine, g)
    r = engine.point(spectra, [], [])
    assert r["nseg"] == 0 and np.all(r["rates"] == 0) and r["ndot_remaining"].shape == (0, 7)
    r = engine.point(spectra, [5, 7], [0, 0])
    assert r["nseg"] == 0 and np.all(r["rates"] == 0) and np.all(r["ndot_boundary"] == 0)
    r = engine.point(spectra, [100], [1], max_pixel_level=1)
    assert r["nseg"] > 0 and np.isclose(r["ndot_boundary"][0, -1], 1.0, rtol=1e-12)
    with pytest.raises(rt.RTB200Error):
        engine.point(spectra, [6 ** 3], [1])                # leaf out of range
    with pytest.raises(rt.RTB200Error):
        engine.point(spectra, [0], [1], max_pixel_level=9)  # beyond the supported pixel levels
    with pytest.raises(rt.RTB200Error):
        engine.point(spectra, [0], [1], dust_approximation=3)
    g1 = W.uniform_grid(1, seed=3, tau_lo=1e-2, tau_hi=1.0, beta24=S24)
    _set(engine, g1)
    r = engine.point(spectra, [0], [4])
    assert r["nseg"] == 12 and np.isclose(r["ndot_boundary"][0, -1], 4.0, rtol=1e-12)   # 12 base rays, one cell each


def _floor_stats(r, o, T):
    """strict per-cell relative error against the libm oracle and the share of cells that need the conditioning floor
    2e-13 * R_r(0) on top of 1e-9 (the noise of exp(interpolated log) that R(d) - R(d+tau) exposes)"""
    worst, need, cells = 0.0, 0, 0
    for i, t in zip(range(6), (0, 2, 1, 3, 5, 4)):
        m = o["rates"][i] != 0
        d = np.abs(r["rates"][i][m] - o["rates"][i][m])
        strict = d / np.abs(o["rates"][i][m])
        assert np.all(d <= TOL * np.abs(o["rates"][i][m]) + 2e-13 * T[t, 0]), (i, float(strict.max()))
        worst = max(worst, float(strict.max()) if strict.size else 0.0)
        need += int(np.sum(strict > TOL)); cells += int(m.sum())
    return worst, need, cells


def test_config1_exact_32cube_h_only(rt, engine, oracle, spectra):
    """BASELINE config 1 as SURVEY.md 8d states it: 32^3, one level, 100 kpc box, HI log-uniform with tau_cell(nu1) in
    [1e-3, 1], HeI = HeII = 0, no dust, one source of weight 1 at the centre of leaf (17,17,17), maxPixelLevel 6,
    numpy default_rng(20240601).  Portable math: bit-identical deposits.  libm oracle: 1e-9 + the floor, with the share
    of cells that need the floor and the worst strict error reported."""
    n = 32
    g = W.uniform_grid(n, seed=20240601, box_kpc=100.0, tau_lo=1e-3, tau_hi=1.0, beta24=S24, helium=False)
    assert not g["HeI"].any() and not g["HeII"].any()
    _set(engine, g)
    leaf = (16 * n + 16) * n + 16                      # (17,17,17), 1-based
    og = _ograd(oracle, g)
    oracle.set_portable_math(True)
    try:
        op = og.point(spectra, [leaf], [1])
    finally:
        oracle.set_portable_math(False)
    ol = og.point(spectra, [leaf], [1])
    assert op["status"] == 0 and ol["status"] == 0
    r = engine.point(spectra, [leaf], [1])             # FAITHFUL + portable exp/log
    assert r["nseg"] == op["nseg"] == ol["nseg"]
    assert 1.5e5 < r["nseg"] < 3e5                     # ~2.1e5 segment updates
    for i in (0, 3):
        assert _cell_err(r["rates"][i], op["rates"][i]) < 1e-12
    assert np.all(r["rates"][[1, 2, 4, 5]] == 0) and np.all(ol["rates"][[1, 2, 4, 5]] == 0)   # H only
    assert np.array_equal(r["highest_pixel_level"], op["highest_pixel_level"])
    assert rel_err(r["ndot_remaining"], op["ndot_remaining"], floor=1e-300) < 1e-12
    assert rel_err(r["ndot_boundary"], op["ndot_boundary"], floor=1e-300) < 1e-12
    T = engine.point_tables(spectra, *_bracket(spectra, g["abun2"][leaf])).reshape(6, -1)
    stats = {}
    for name, mode in (("faithful", rt.MATH_FAITHFUL), ("fast", rt.MATH_FAST)):
        engine.set_math(mode)
        rr = engine.point(spectra, [leaf], [1])
        assert rr["nseg"] == ol["nseg"]
        worst, need, cells = _floor_stats(rr, ol, T)
        stats[name] = (worst, need, cells)
        assert rel_err(rr["ndot_remaining"], ol["ndot_remaining"], floor=1e-300) < 1e-11
        assert np.allclose(rr["rates"].sum(axis=1), ol["rates"].sum(axis=1), rtol=1e-11, atol=0)
    print("config 1 vs libm oracle (worst strict rel err, cells above 1e-9, cells):", stats)
    for worst, need, cells in stats.values():
        assert need <= 0.02 * cells                    # the floor is an exception, not the rule


def test_escaping_spectrum_and_highest_pixel_level(rt, engine, oracle, spectra):
    """ndotSpectrum(300) is only filled where a segment straddles the last output radius, 100 kpc
    (equiSources.f90:3206-3224): a 400 kpc box with sources near the centre makes every leaf ray cross it.  Also
    ndotDust, and highestPixelLevel (:3316), on a uniform and a nested grid, with and without dust."""
    cases = []
    g = W.uniform_grid(16, seed=31, box_kpc=400.0, tau_lo=1e-3, tau_hi=0.3, beta24=S24)
    g["abun2"] = np.random.default_rng(3).uniform(1e-3, 4e-2, 16 ** 3)
    cases.append((g, [(8 * 16 + 7) * 16 + 8, (6 * 16 + 9) * 16 + 7], [1, 2]))
    g = W.nested_grid(8, 2, W.central_box_refine(0.3, 0.7, levels=2), seed=32, box_kpc=400.0, tau_lo=1e-3, tau_hi=0.3,
                      beta24=S24)
    cx, cy, cz = g["centres"]
    cases.append((g, [int(np.argmin((cx - 0.52) ** 2 + (cy - 0.47) ** 2 + (cz - 0.5) ** 2))], [3]))
    for g, srcs, wts in cases:
        _set(engine, g)
        og = _ograd(oracle, g)
        for dust in (0, 1):
            for maxlev in (6, 3):
                oracle.set_portable_math(True)
                try:
                    op = og.point(spectra, srcs, wts, dust_approximation=dust, max_pixel_level=maxlev)
                finally:
                    oracle.set_portable_math(False)
                ol = og.point(spectra, srcs, wts, dust_approximation=dust, max_pixel_level=maxlev)
                engine.set_math(rt.MATH_FAITHFUL)
                r = engine.point(spectra, srcs, wts, dust_approximation=dust, max_pixel_level=maxlev)
                assert r["nseg"] == op["nseg"]
                assert op["ndot_spectrum"].min() > 0           # every energy bin received photons
                assert rel_err(r["ndot_spectrum"], op["ndot_spectrum"], floor=1e-300) < 1e-12
                assert rel_err(r["ndot_dust"], op["ndot_dust"], floor=1e-300) < 1e-12
                assert rel_err(r["ndot_spectrum"], ol["ndot_spectrum"], floor=1e-300) < 1e-11   # libm oracle
                assert np.array_equal(r["highest_pixel_level"], op["highest_pixel_level"])
                assert np.all(r["highest_pixel_level"] == maxlev)   # rays of this size always reach the last level
                engine.set_math(rt.MATH_FAST)
                rf = engine.point(spectra, srcs, wts, dust_approximation=dust, max_pixel_level=maxlev)
                assert rel_err(rf["ndot_spectrum"], ol["ndot_spectrum"], floor=1e-300) < 1e-11
                assert np.array_equal(rf["highest_pixel_level"], ol["highest_pixel_level"])
                # no dust: the dust term is exp(0) per ray -> ndotDust = photons that reach 100 kpc unattenuated by dust
                if dust == 0:
                    assert np.allclose(r["ndot_dust"], np.array(wts, dtype=float) - ol["ndot_boundary"][:, -1], rtol=1e-12)
    # a source that never splits: one pixel level
    r = engine.point(spectra, srcs, wts, max_pixel_level=1)
    assert np.all(r["highest_pixel_level"] == 0)


def test_golden_spectrum_fixture(rt, engine, spectra):
    """committed libm-oracle output for the escaping spectrum (tests/golden/point_spectrum_12.npz)"""
    import os
    from conftest import ROOT
    f = np.load(os.path.join(ROOT, "tests", "golden", "point_spectrum_12.npz"))
    g = W.uniform_grid(12, seed=int(f["seed"]), box_kpc=float(f["box_kpc"]), tau_lo=1e-3, tau_hi=0.3, beta24=S24)
    _set(engine, g)
    for mode in (rt.MATH_FAITHFUL, rt.MATH_FAST):
        engine.set_math(mode)
        r = engine.point(spectra, f["src"], f["wt"], dust_approximation=int(f["dust"]))
        assert r["nseg"] == int(f["nseg"])
        assert rel_err(r["ndot_spectrum"], f["ndot_spectrum"], floor=1e-300) < 1e-11
        assert rel_err(r["ndot_dust"], f["ndot_dust"], floor=1e-300) < 1e-11
        assert np.array_equal(r["highest_pixel_level"], f["highest_pixel_level"])


def test_planned_deposition_atomic_free_and_reproducible(rt, engine, portable, spectra):
    """set_tuning(point_deposit=2): the (leaf, ray, segment) sort is done once per (grid, sources); later passes write
    every deposit to its cached leaf-ordered slot and add up each leaf's run -- no atomics, no sort.  Bit-identical to
    the sort-every-pass mode (same summation order), from call to call, and after the species change; with dust (rays
    may end early) the mode falls back to sorting every pass."""
    n = 12
    g = W.nested_grid(n, 1, W.central_box_refine(0.25, 0.75, levels=1), seed=21, tau_lo=1e-2, tau_hi=1.0, beta24=S24)
    _set(engine, g)
    rng = np.random.default_rng(8)
    leaves = rng.choice(g["level"].size, 70, replace=False).astype(np.int32)      # more than one batch of 64 sources
    wts = rng.integers(0, 3, 70).astype(np.int32)
    engine.set_tuning(point_deposit=0)
    atom = engine.point(spectra, leaves, wts)
    engine.set_tuning(point_deposit=1)
    srt = engine.point(spectra, leaves, wts)
    engine.set_tuning(point_deposit=2)
    p1 = engine.point(spectra, leaves, wts)          # plan pass
    p2 = engine.point(spectra, leaves, wts)          # planned pass
    p3 = engine.point(spectra, leaves, wts)
    assert p1["nseg"] == p2["nseg"] == atom["nseg"]
    for r in (p1, p2, p3):
        for i in range(6):
            assert _cell_err(r["rates"][i], atom["rates"][i]) < 1e-12
    assert np.array_equal(p2["rates"], p3["rates"])
    assert np.array_equal(p1["rates"], p2["rates"])
    assert np.array_equal(p2["ndot_boundary"], atom["ndot_boundary"])
    # same sources, other densities: the plan stays valid
    engine.update_species(HI=g["HI"] * 0.3, HeI=g["HeI"] * 2.0)
    q2 = engine.point(spectra, leaves, wts)
    engine.set_tuning(point_deposit=0)
    q0 = engine.point(spectra, leaves, wts)
    for i in range(6):
        assert _cell_err(q2["rates"][i], q0["rates"][i]) < 1e-12
    o = _ograd(portable, dict(g, HI=g["HI"] * 0.3, HeI=g["HeI"] * 2.0)).point(spectra, leaves, wts)
    for i in range(6):
        assert _cell_err(q2["rates"][i], o["rates"][i]) < 1e-12
    # other sources: a new plan; accumulation into existing rates
    engine.set_tuning(point_deposit=2)
    r1 = engine.point(spectra, leaves[:5], wts[:5] + 1)
    r2 = engine.point(spectra, leaves[:5], wts[:5] + 1, rates=r1["rates"])
    assert rel_err(r2["rates"], 2 * r1["rates"], floor=1e-300) < 1e-12
    # dust: falls back to the sort-every-pass mode, still reproducible
    d1 = engine.point(spectra, leaves[:5], wts[:5] + 1, dust_approximation=2)
    d2 = engine.point(spectra, leaves[:5], wts[:5] + 1, dust_approximation=2)
    assert np.array_equal(d1["rates"], d2["rates"])
    engine.set_tuning(point_deposit=0)
