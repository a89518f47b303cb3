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


class _Cell:
    __slots__ = ("s", "kids", "level")

    def __init__(self, level):
        self.s = dict(tgas=0., rho=0., HI=0., HeI=0., HeII=0., abun2=0.)
        self.kids, self.level = None, level


def _reference_build(levels, metals):
    """pointer-tree restatement of equiSources.f90:425-618 + placeCellProjectWithVelocity (:1870-1974), one cell at a time"""
    f32 = np.float32
    psi, mp, mn = float(f32(0.76)), float(f32(1.6726231e-24)), float(f32(1.67492728e-24))
    mhe = 2.0 * (mp + mn)
    n1 = levels[0]["pos"].shape[0]
    nx = round(n1 ** (1 / 3))
    p1 = levels[0]["pos"].astype(f32).astype(float)
    lo, hi = p1.min(0), p1.max(0)
    mid, half = 0.5 * (lo + hi), 0.5 * (hi - lo) * float(nx) / float(nx - 1)
    a, b = mid - half, mid + half
    base = [[[_Cell(0) for _ in range(nx)] for _ in range(nx)] for _ in range(nx)]

    def place(cell, level, x, y, z, rec):
        if level > 1:
            if cell.kids is None:
                cell.kids = [[[_Cell(cell.level + 1) for _ in range(2)] for _ in range(2)] for _ in range(2)]
                for i in range(2):
                    for j in range(2):
                        for k in range(2):
                            c = cell.kids[i][j][k]
                            c.s.update(cell.s)
                            c.s["abun2"] = 0.0
            i, x = (0, 2 * x) if x < 0.5 else (1, 2 * x - 1)
            j, y = (0, 2 * y) if y < 0.5 else (1, 2 * y - 1)
            k, z = (0, 2 * z) if z < 0.5 else (1, 2 * z - 1)
            place(cell.kids[i][j][k], level - 1, x, y, z, rec)
        else:
            nh = 10.0 ** float(rec["lnH"])
            rho = nh * mp / psi
            cell.s.update(tgas=10.0 ** float(rec["lT"]), rho=rho, HI=nh * 10.0 ** float(rec["lx"]),
                          HeI=(1 - psi) * rho / mhe, HeII=0.0,
                          abun2=float(rec["abun"][1]) if metals else float(f32(0.02)))

    smoothed = None
    if metals:
        # smoothing level 1 abundancies (equiSources.f90:527-578), the reference's scatter loops as they stand
        lv = levels[0]
        pn = ((lv["pos"].astype(f32).astype(float) - a) / (b - a)).astype(f32).astype(float)
        uni = np.zeros((nx, nx, nx))
        cellIdx = [(int(x * nx), int(y * nx), int(z * nx)) for x, y, z in pn]
        for c, ijk in enumerate(cellIdx):
            uni[ijk] = float(f32(lv["abun"][c][1]))
        for _ in range(2):
            for ax in range(3):
                tmp = np.zeros((nx, nx, nx))
                for i in range(nx):
                    for j in range(nx):
                        for k in range(nx):
                            tmp[i, j, k] = tmp[i, j, k] + 0.5 * uni[i, j, k]
                            m = [i, j, k]
                            if m[ax] > 0:
                                q = list(m); q[ax] -= 1
                                tmp[tuple(q)] = tmp[tuple(q)] + 0.25 * uni[i, j, k]
                            if m[ax] < nx - 1:
                                q = list(m); q[ax] += 1
                                tmp[tuple(q)] = tmp[tuple(q)] + 0.25 * uni[i, j, k]
                uni = tmp
        smoothed = [f32(uni[ijk]) for ijk in cellIdx]

    for l, lv in enumerate(levels):
        pn = ((lv["pos"].astype(f32).astype(float) - a) / (b - a)).astype(f32).astype(float)
        for c in range(pn.shape[0]):
            x0, y0, z0 = pn[c]
            i0, j0, k0 = int(x0 * nx), int(y0 * nx), int(z0 * nx)
            rec = dict(lT=f32(lv["lT"][c]), lnH=f32(lv["lnH"][c]), lx=f32(lv["lx"][c]))
            if metals:
                rec["abun"] = lv["abun"][c].astype(f32)
                if l == 0:
                    rec["abun"][1] = smoothed[c]
            place(base[i0][j0][k0], l + 1, x0 * float(nx) - i0, y0 * float(nx) - j0, z0 * float(nx) - k0, rec)
    out = []

    def walk(cell):
        if cell.kids is None:
            out.append((cell.level, cell.s))
        else:
            for i in range(2):
                for j in range(2):
                    for k in range(2):
                        walk(cell.kids[i][j][k])

    for i in range(nx):
        for j in range(nx):
            for k in range(nx):
                walk(base[i][j][k])
    return nx, out


@pytest.mark.parametrize("metals", [False, True])
def test_octree_from_level_lists(metals, tmp_path, oracle):
    rng = np.random.default_rng(5)
    nx = 4
    g = (np.arange(nx) + 0.5) / nx * 80.0 - 40.0            # kpc
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")

    def lv(pos):
        n = pos.shape[0]
        d = dict(pos=pos, lT=rng.uniform(3, 5, n), lnH=rng.uniform(-4, 0, n), lx=rng.uniform(-4, 0, n))
        if metals:
            d["abun"] = rng.uniform(0, 0.05, (n, 4))
        return d

    levels = [lv(np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1)),
              lv(rng.uniform(-12, 12, (40, 3))),                 # level 2 cells near the centre (some share a parent)
              lv(rng.uniform(-6, 6, (30, 3)))]                    # level 3 cells: parents partly created by inheritance
    p = tmp_path / ("grid_met.dat" if metals else "grid.dat")
    F.write_grid_dat(p, levels)
    back = F.read_grid_dat(p, metals=metals)
    kw = F.build_leaves(back, metals=metals)
    nref, leaves = _reference_build(back, metals)
    assert kw["nx"] == nref and kw["level"].size == len(leaves)
    assert np.array_equal(kw["level"], np.array([l for l, _ in leaves], dtype=np.int8))
    for name in ("HI", "HeI", "HeII", "rho", "abun2", "tgas"):
        assert np.array_equal(kw[name], np.array([s[name] for _, s in leaves])), name
    assert np.isclose(kw["box_size"], 80.0 * F._KPC, rtol=1e-6)
    tgas = kw.pop("tgas")
    og = oracle.OracleGrid(kw["nx"], kw["level"], kw["HI"], kw["HeI"], kw["HeII"], kw["rho"], kw["abun2"], kw["box_size"])
    assert og.nleaf == tgas.size                                  # a consistent pre-order octree


def test_hdf4_stub_container_round_trip(tmp_path):
    """oracle/ref_harness/hdf4_stub.c (the HDF4 stand-in of the gfortran recipe) against formats.write_sd_container /
    read_sd_container: datasets written by Python are read through the Fortran-callable entry points and back"""
    import ctypes as C
    import os
    import subprocess
    from conftest import ROOT
    so = str(tmp_path / "libhdf4stub.so")
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "oracle", "ref_harness", "hdf4_stub.c")])
    L = C.CDLL(so)
    rng = np.random.default_rng(3)
    pos = rng.normal(size=(27, 3)).astype("<f4")
    lT = rng.normal(size=27).astype("<f4")
    path = str(tmp_path / "grid_met.h4")
    F.write_sd_container(path, [("nlevels", np.array([1], dtype="<i4")), ("pos1", pos), ("lT1", lT)])
    back = F.read_sd_container(path)
    assert [n for n, _ in back] == ["nlevels", "pos1", "lT1"] and np.array_equal(back[1][1], pos)
    i = lambda v: C.byref(C.c_int(v))
    name = (path + "   ").encode()                        # Fortran strings are blank padded, length passed by value
    sd = L.sfstart_(name, i(1), C.c_long(len(name)))
    assert sd > 0
    nds, nat = C.c_int(0), C.c_int(0)
    assert L.sffinfo_(i(sd), C.byref(nds), C.byref(nat)) == 0 and nds.value == 3
    sds = L.sfselect_(i(sd), i(1))
    nm = C.create_string_buffer(64)
    rank, typ, na = C.c_int(0), C.c_int(0), C.c_int(0)
    dims = (C.c_int * 4)()
    assert L.sfginfo_(i(sds), nm, C.byref(rank), dims, C.byref(typ), C.byref(na), C.c_long(64)) == 0
    assert rank.value == 2 and dims[0] == 27 and dims[1] == 3 and typ.value == 5 and nm.raw.startswith(b"pos1 ")
    out = np.zeros((27, 3), dtype="<f4", order="F")      # a Fortran array pos(27,3)
    start, stride, edges = (C.c_int * 4)(0, 0, 0, 0), (C.c_int * 4)(1, 1, 1, 1), (C.c_int * 4)(27, 3, 0, 0)
    assert L.sfrdata_(i(sds), start, stride, edges, out.ctypes.data_as(C.c_void_p)) == 0
    assert np.array_equal(out, pos)
    assert L.sfendacc_(i(sds)) == 0 and L.sfend_(i(sd)) == 0
    # writing, as writeIonization does (equiSources.f90:4843-4905)
    path2 = (str(tmp_path / "cellArray0001.h4")).encode()
    sd = L.sfstart_(path2, i(4), C.c_long(len(path2)))
    lev = np.arange(10, dtype="<i4")
    e1 = (C.c_int * 4)(10, 0, 0, 0)
    sds = L.sfcreate_(i(sd), b"level", i(24), i(1), e1, C.c_long(5))
    assert L.sfwdata_(i(sds), start, stride, e1, lev.ctypes.data_as(C.c_void_p)) == 0
    hi = rng.normal(size=10).astype("<f4")
    sds = L.sfcreate_(i(sd), b"HI", i(5), i(1), e1, C.c_long(2))
    assert L.sfwdata_(i(sds), start, stride, e1, hi.ctypes.data_as(C.c_void_p)) == 0
    assert L.sfend_(i(sd)) == 0
    back = F.read_sd_container(path2.decode())
    assert back[0][0] == "level" and np.array_equal(back[0][1], lev) and np.array_equal(back[1][1], hi)


def _host_backend_octree(tmp_path):
    """csrc/octree_build.cu compiled with thrust's HOST backend (g++, THRUST_DEVICE_SYSTEM_CPP): the same source the
    product builds for the GPU, runnable without one.  A test artefact: its entry points carry another prefix."""
    import ctypes as C
    import os
    import subprocess
    from conftest import ROOT
    so = str(tmp_path / "liboctree_hostcheck.so")
    inc = "/usr/local/cuda/include"
    if not os.path.exists(os.path.join(inc, "thrust", "sort.h")):
        pytest.skip("thrust headers not found")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-x", "c++", "-DTHRUST_DEVICE_SYSTEM=THRUST_DEVICE_SYSTEM_CPP", "-I", inc,
                           "-fPIC", "-ffp-contract=off", "-shared", "-o", so,
                           os.path.join(ROOT, "radiativetransfer_b200", "csrc", "octree_build.cu")])
    return C.CDLL(so)


def _random_levels(rng, nx, metals, counts=(40, 30), spans=(12.0, 6.0), duplicates=True):
    g = (np.arange(nx) + 0.5) / nx * 80.0 - 40.0            # kpc
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")

    def lv(pos):
        n = pos.shape[0]
        d = dict(pos=pos, lT=rng.uniform(3, 5, n), lnH=rng.uniform(-4, 0, n), lx=rng.uniform(-4, 0, n))
        if metals:
            d["abun"] = rng.uniform(0, 0.05, (n, 4))
        return d

    levels = [lv(np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1))]
    for c, s in zip(counts, spans):
        pos = rng.uniform(-s, s, (c, 3))
        if duplicates and c > 4:
            pos[-2:] = pos[:2] + 1e-4          # two cells that land in the same node: the later one wins
        levels.append(lv(pos))
    return levels


@pytest.mark.parametrize("metals", [False, True])
def test_octree_build_sorted_key_sets_equal_sequential_insertion(metals, tmp_path):
    """the device formulation of the octree build (sorted key sets, csrc/octree_build.cu), run through thrust's host
    backend, against formats.build_leaves (itself checked against the pointer-tree restatement above): same leaves,
    same order, same bits"""
    L = _host_backend_octree(tmp_path)
    rng = np.random.default_rng(11)
    for nx, counts, spans in ((4, (40, 30), (12.0, 6.0)), (3, (25,), (20.0,)), (5, (60, 50, 40), (16.0, 9.0, 5.0)), (2, (), ())):
        levels = _random_levels(rng, nx, metals, counts, spans)
        for lv in levels:                      # what read_grid_dat hands over: real*4 lists
            for k in lv:
                lv[k] = np.asarray(lv[k], dtype=np.float32)
        ref = F.build_leaves(levels, metals=metals)
        got = F.build_leaves_device(levels, metals=metals, _library=L, _prefix="rtb200_hostcheck_octree_")
        assert got["nx"] == ref["nx"] and got["box_size"] == ref["box_size"]
        assert np.array_equal(got["level"], ref["level"])
        for name in ("HI", "HeI", "HeII", "rho", "abun2", "tgas"):
            assert np.array_equal(got[name], ref[name]), (nx, name)
    # a level-1 list that is not a cube, and a cell outside the box, are refused
    bad = _random_levels(rng, 3, metals, (), ())
    bad[0] = {k: v[:-1] for k, v in bad[0].items()}
    with pytest.raises(RuntimeError):
        F.build_leaves_device(bad, metals=metals, _library=L, _prefix="rtb200_hostcheck_octree_")
    out = _random_levels(rng, 3, metals, (5,), (10.0,))
    out[1]["pos"][0] = [500.0, 0.0, 0.0]
    with pytest.raises(RuntimeError):
        F.build_leaves_device(out, metals=metals, _library=L, _prefix="rtb200_hostcheck_octree_")


@pytest.mark.gpu
@pytest.mark.parametrize("metals", [False, True])
def test_octree_build_on_device(metals, build_product):
    """rtb200_octree_build on the GPU against formats.build_leaves: tree, leaf order and every integer bit-exact; the
    state differs only by the device's pow(10, x) (within 2 ulp of the C library's; HI is a product of two) -- 1e-15
    relative; and the result feeds
    rtb200_grid_set"""
    rng = np.random.default_rng(12)
    import radiativetransfer_b200 as rt
    for nx, counts, spans in ((4, (40, 30), (12.0, 6.0)), (16, (3000, 2000, 800), (20.0, 10.0, 4.0))):
        levels = _random_levels(rng, nx, metals, counts, spans)
        for lv in levels:
            for k in lv:
                lv[k] = np.asarray(lv[k], dtype=np.float32)
        ref = F.build_leaves(levels, metals=metals)
        got = F.build_leaves_device(levels, metals=metals)
        assert got["nx"] == ref["nx"] and got["box_size"] == ref["box_size"]
        assert np.array_equal(got["level"], ref["level"])
        assert np.array_equal(got["HeII"], ref["HeII"]) and np.array_equal(got["abun2"], ref["abun2"])
        for name in ("HI", "HeI", "rho", "tgas"):
            m = ref[name] != 0
            assert np.array_equal(got[name] == 0, ~m)
            assert np.max(np.abs(got[name][m] / ref[name][m] - 1.0)) < 1e-15, name
        t = rt.Transport(device=0)
        tg = got.pop("tgas")
        t.set_grid(**got)
        assert t.nleaf == tg.size
        t.close()


def test_level_sequences_are_validated():
    """count_base_cells walks the pre-order like readCellArray.f90:154-187: an octet that closes early, a refined cell
    without all 8 children and a volume count that would overflow int64 are refused"""
    ok = np.array([1] * 3 + [2] * 8 + [1] * 4 + [0] * 7, dtype=np.int8)
    assert F.count_base_cells(ok) == 8
    for bad in ([0, 1, 1, 1, 1, 1, 1, 1, 0], [1] * 7 + [0] * 8, [0, 2] + [1] * 7, [1, 1, 1, 1] + [2] * 7 + [1] * 4, [1] * 8 + [2] * 7):
        with pytest.raises(ValueError):
            F.count_base_cells(np.array(bad, dtype=np.int8))
    with pytest.raises(ValueError):
        F.count_base_cells(np.array([-1, 0], dtype=np.int8))
    deep = np.array([20] * 8, dtype=np.int8)                  # 3 * 20 + bits(8 leaves) >= 63
    with pytest.raises(ValueError):
        F.leaf_centres(1, deep)
