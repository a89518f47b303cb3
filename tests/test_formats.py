"""HDF-free file formats of the unchanged tools: Fortran unformatted `cellArray.dat`
(hdf42bin.f90:208-218) and grid `.dat` (bin2hdf4.f90:43-83)."""
import numpy as np
import pytest

from radiativetransfer_b200 import formats as F, workloads as W


def test_record_markers(tmp_path):
    p = tmp_path / "r.dat"
    F.write_records(p, [np.arange(3, dtype="<i4"), np.array([1.5], dtype="<f4")])
    raw = p.read_bytes()
    assert raw[:4] == (12).to_bytes(4, "little") and raw[16:20] == (12).to_bytes(4, "little")   # gfortran framing
    assert len(raw) == 4 + 12 + 4 + 4 + 4 + 4
    a, b = F.read_records(p)
    assert np.array_equal(np.frombuffer(a, "<i4"), [0, 1, 2]) and np.frombuffer(b, "<f4")[0] == 1.5
    p.write_bytes(raw[:-2])
    with pytest.raises(ValueError):
        F.read_records(p)


def test_cell_array_round_trip_and_tree_from_levels(tmp_path, oracle):
    g = W.nested_grid(4, 2, W.central_box_refine(0.25, 0.75, levels=2), seed=3)
    x, y, z = F.leaf_centres(4, g["level"])
    cx, cy, cz = g["centres"]
    assert np.allclose(x, cx) and np.allclose(y, cy) and np.allclose(z, cz)   # the leaf order IS the writeCell pre-order
    p = tmp_path / "cellArray0001.dat"
    T = np.full(g["level"].size, 1.0e4)
    F.write_cell_array_dat(p, g["level"], x, y, z, g["HI"], g["HeI"], g["HeII"], T, g["rho"])
    c = F.read_cell_array_dat(p)
    assert c["nx"] == 4 and c["nleaf"] == g["level"].size
    assert np.array_equal(c["level"], g["level"])
    assert np.array_equal(c["HI"], g["HI"].astype(np.float32))                 # the files hold single precision
    kw = F.transport_inputs(c, g["box_size"])
    og = oracle.OracleGrid(kw["nx"], kw["level"], kw["HI"], kw["HeI"], kw["HeII"], kw["rho"], kw["abun2"], kw["box_size"])
    assert og.nleaf == c["nleaf"]                                              # same tree from `level` alone
    bad = g["level"].copy(); bad[0] = 1
    with pytest.raises(ValueError):
        F.count_base_cells(bad)


def test_grid_dat_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    levels = []
    for n in (8, 5):
        levels.append(dict(pos=rng.random((n, 3)), lT=rng.random(n), lnH=rng.random(n), lx=rng.random(n),
                           abun=rng.random((n, 4)), vel=rng.random((n, 3))))
    p = tmp_path / "grid_met_vel.dat"
    F.write_grid_dat(p, levels)
    back = F.read_grid_dat(p, metals=True, kinematics=True)
    assert len(back) == 2
    for a, b in zip(levels, back):
        for k in a:
            assert np.array_equal(np.asarray(a[k], dtype=np.float32), b[k])
    with pytest.raises((ValueError, StopIteration)):
        F.read_grid_dat(p, metals=False, kinematics=False)
