"""Known-answer tests pinning the CPU oracle (the reference ships no golden vectors: parity is unpinned by the
reference itself, SURVEY.md section 4 lists the analytic checks used here)."""
import numpy as np
import pytest

from radiativetransfer_b200 import workloads as W


def test_all_24_zones_populated(oracle):
    zones = {}
    for r in range(192):
        st, iz, phi, th = oracle.direction(3, r)
        assert st == 0
        assert 0 < phi < np.pi / 2 and 0 < th < np.pi / 2
        zones[iz] = zones.get(iz, 0) + 1
    assert sorted(zones) == list(range(1, 25))
    assert sum(zones.values()) == 192


def test_healpix_level1_pixel_centres(oracle):
    # un-rotated HEALPix nside=1 centres have z in {2/3, 0, -2/3}; the fixed tilt keeps |direction| = 1 and the
    # 12 directions distinct
    vec = []
    for p in range(12):
        st, phi, th = oracle.pix2ang_nest(1, p)
        assert st == 0
        vec.append([np.cos(phi) * np.cos(th), np.sin(phi) * np.cos(th), np.sin(th)])
    vec = np.array(vec)
    assert np.allclose(np.linalg.norm(vec, axis=1), 1.0)
    assert np.allclose(vec.sum(axis=0), 0.0, atol=1e-6)  # pi is single precision in the reference
    d = np.linalg.norm(vec[:, None] - vec[None], axis=2) + np.eye(12)
    assert d.min() > 0.5


def test_rotate_indices_is_a_bijection(oracle):
    n = 4
    for z in range(1, 25):
        seen = set()
        for i in range(1, n + 1):
            for j in range(1, n + 1):
                for k in range(1, n + 1):
                    seen.add(oracle.rotate_indices(i, j, k, n, n, n, z))
        assert len(seen) == n ** 3
        assert all(1 <= c <= n for t in seen for c in t)


@pytest.mark.parametrize("n", [1, 5, 16])
def test_zero_opacity_gives_background(oracle, uvbg, n):
    # kappa = 0 -> every segment takes the arithmetic branch -> J = uvb * 192 * float(1/192) to a few ulp
    g = oracle.OracleGrid(n, np.zeros(n ** 3, dtype=np.int8), np.zeros(n ** 3), box_size=3e23)
    o = g.diffuse(uvbg["uvb"], uvbg["beta"])
    assert o["status"] == 0
    w = float(np.float32(1) / np.float32(192))
    for gidx in range(3):
        assert np.allclose(o["J"][gidx], uvbg["uvb"][gidx] * 192 * w, rtol=1e-14, atol=0)


def test_zero_opacity_amr(oracle, uvbg):
    g = W.nested_grid(8, 2, W.central_box_refine(0.25, 0.75, levels=2), seed=2)
    og = oracle.OracleGrid(8, g["level"], np.zeros(g["level"].size), box_size=g["box_size"])
    o = og.diffuse(uvbg["uvb"], uvbg["beta"])
    assert o["status"] == 0
    w = float(np.float32(1) / np.float32(192))
    assert np.allclose(o["J"][0], uvbg["uvb"][0] * 192 * w, rtol=1e-14, atol=0)


def test_pattern_lengths_sum_to_chord(oracle, uvbg):
    # per layer: len_xy + len_xz + len_yz = 1/sin(theta) (what checkPattern prints, transportRoutinesModule.f90:245)
    n = 32
    g = oracle.OracleGrid(n, np.zeros(n ** 3, dtype=np.int8), np.ones(n ** 3), box_size=1.0)
    for ray in (0, 17, 100, 191):
        o = g.diffuse(uvbg["uvb"], uvbg["beta"], ray_begin=ray, ray_end=ray + 1, trace_ray=ray)
        p, th = o["pattern"], o["angles"][1]
        assert np.allclose(p[:, 2] + p[:, 5] + p[:, 8], 1 / np.sin(th), rtol=1e-12)
        assert (p[:, [0, 1, 3, 4, 6, 7]] >= 0).all() and (p[:, [0, 1, 3, 4, 6, 7]] <= 1).all()
        nseg = 1 + (p[:, 5] > 0) + (p[:, 8] > 0)
        assert o["nseg"] == int(nseg.sum()) * n * n


def test_single_cell_segment_math(oracle, uvbg):
    # n = 1: no neighbours, every segment starts from the background.  Independent numpy evaluation.
    kappa_HI = 3.0e-3
    g = oracle.OracleGrid(1, np.zeros(1, dtype=np.int8), np.array([kappa_HI]), box_size=2.0e20)
    beta = uvbg["beta"]
    uvb = uvbg["uvb"]
    w = float(np.float32(1) / np.float32(192))
    J = np.zeros(3)
    for ray in range(192):
        o = g.diffuse(uvb, beta, ray_begin=ray, ray_end=ray + 1, trace_ray=ray)
        p = o["pattern"][0]
        lens = [l for l in (p[2], p[5], p[8]) if l > 0]
        kap = np.array([kappa_HI * beta[0, 0], kappa_HI * beta[1, 0], kappa_HI * beta[2, 0]])
        acc = np.zeros(3)
        for l in lens:
            tau = kap * 2.0e20 * l
            acc += uvb * (-np.expm1(-tau)) / tau
        J += acc / len(lens) * w
    full = g.diffuse(uvb, beta)
    assert np.allclose(full["J"][:, 0], J, rtol=1e-10)


def test_uniform_opacity_first_layer_attenuation(oracle, uvbg):
    # uniform kappa: a cell on the inflow corner sees only background-fed segments
    n = 6
    HI = np.full(n ** 3, 1.0e-3)
    g = oracle.OracleGrid(n, np.zeros(n ** 3, dtype=np.int8), HI, box_size=6.0e20)
    uvb, beta = uvbg["uvb"], uvbg["beta"]
    ray = 3
    o = g.diffuse(uvb, beta, ray_begin=ray, ray_end=ray + 1, trace_ray=ray)
    nb = o["nb"]
    corner = np.where((nb[0] == -1) & (nb[1] < 0) & (nb[2] < 0) & ((nb[1] == -1) | (nb[1] == -2)) &
                      ((nb[2] == -1) | (nb[2] == -2)))[0]
    assert corner.size >= 1
    p = o["pattern"][0]
    lens = [l for l in (p[2], p[5], p[8]) if l > 0]
    tau = HI[0] * beta[0, 0] * 1.0e20 * np.array(lens)
    expect = np.mean(uvb[0] * (-np.expm1(-tau)) / tau) * float(np.float32(1) / np.float32(192))
    assert np.allclose(o["J"][0][corner], expect, rtol=1e-10)


def test_direction_shards_add_up(oracle, uvbg):
    g = W.uniform_grid(8, seed=4)
    og = oracle.OracleGrid(8, g["level"], g["HI"], g["HeI"], g["HeII"], box_size=g["box_size"])
    full = og.diffuse(uvbg["uvb"], uvbg["beta"])["J"]
    a = og.diffuse(uvbg["uvb"], uvbg["beta"], ray_begin=0, ray_end=100)["J"]
    b = og.diffuse(uvbg["uvb"], uvbg["beta"], ray_begin=100, ray_end=192)["J"]
    assert np.allclose(a + b, full, rtol=1e-13, atol=0)


def test_level_array_errors(oracle):
    with pytest.raises(RuntimeError):
        oracle.OracleGrid(2, np.array([0] * 7 + [1] * 7, dtype=np.int8), np.ones(14))  # incomplete octet


def test_reference_harness_pieces(tmp_path):
    """oracle/ref_harness (the gfortran recipe that would pin the oracle against the reference itself): what can run
    without a Fortran compiler -- the dump reader / comparer on a dump the oracle writes in the harness layout, and the
    case generator's files"""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    h = os.path.join(ROOT, "oracle", "ref_harness")
    r = subprocess.run([sys.executable, os.path.join(h, "compare.py"), "--selftest", str(tmp_path / "st")], capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0 and "PINNED" in r.stdout, r.stdout + r.stderr
    r = subprocess.run([sys.executable, os.path.join(h, "make_case.py"), str(tmp_path / "run"), "--case", "nested"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    from radiativetransfer_b200 import formats as F
    ds = F.read_sd_container(str(tmp_path / "run" / "case_met.h4"))
    assert ds[0][0] == "nlevels" and int(ds[0][1][0]) == 3
    assert ds[1][1].shape == (512, 3) and len(ds) == 1 + 3 * 5          # 8^3 level-1 cells; pos, lT, lnH, lx, abun per level
    # the spectrum table parses with the driver's fixed columns (equiSources.f90:861-879)
    lines = open(tmp_path / "run" / "model41-salpeter-burst34" / "spectrum.out").read().splitlines()
    data = [l for l in lines if len(l) >= 41 and l[1:10] != "TIME [YR]" and l[1:6] != "MODEL" and l.strip()]
    assert len(data) == 37 * 1221
    assert float(data[0][1:13]) == 1.0e6 and abs(float(data[0][13:28]) - 91.0) < 1e-4 and float(data[0][28:41]) > 0
