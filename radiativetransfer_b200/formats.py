"""HDF-free on-disk formats either side of the hot path.

The reference driver reads its grid and writes its cell arrays through HDF4 (not available here); its own converter
tools define equivalent Fortran *unformatted sequential* files, which this module reads and writes with numpy:

* `cellArrayNNNN.dat` -- hdf42bin.f90:208-218 / convertFormats.f90:263-273: nine records `level` (int32), `x, y, z, HI,
  HeI, HeII, T, rho` (float32), one value per leaf in the leaf pre-order of `writeCell` (equiSources.f90:4044-4079,
  4830-4836) -- the order `rtb200_grid_set` takes.
* `<grid>.dat` -- bin2hdf4.f90:43-83: `nlevels`, then per level `ncell`, `pos(:,1..3)`, `lT`, `lnH`, `lx`
  [, `abun(:,1..4)`][, `vel(:,1..3)`], all float32.

gfortran writes every record as <int32 length> payload <int32 length>, little-endian on x86-64.
"""
import numpy as np

CELL_ARRAY_FIELDS = ("level", "x", "y", "z", "HI", "HeI", "HeII", "T", "rho")


def write_records(path, arrays):
    """Fortran unformatted sequential file: one record per array (4-byte markers)"""
    with open(path, "wb") as f:
        for a in arrays:
            b = np.ascontiguousarray(a).tobytes()
            if len(b) >= 2 ** 31:
                raise ValueError("record longer than 2 GiB: gfortran would split it into sub-records")
            m = np.int32(len(b)).tobytes()
            f.write(m); f.write(b); f.write(m)


def read_records(path):
    """list of raw payloads (bytes) of a Fortran unformatted sequential file"""
    out = []
    with open(path, "rb") as f:
        while True:
            h = f.read(4)
            if not h:
                break
            if len(h) != 4:
                raise ValueError("truncated record marker")
            n = int(np.frombuffer(h, dtype="<i4")[0])
            if n < 0:
                raise ValueError("sub-records (records of 2 GiB or more) are not supported")
            b = f.read(n)
            t = f.read(4)
            if len(b) != n or len(t) != 4 or int(np.frombuffer(t, dtype="<i4")[0]) != n:
                raise ValueError("corrupt record: leading and trailing markers differ")
            out.append(b)
    return out


def count_base_cells(level):
    """number of base cells of a leaf-level array in pre-order (the tree is implied by `level` alone, as
    readCellArray.f90:154-187 rebuilds it); raises on an inconsistent sequence ('error in levels')"""
    level = np.asarray(level, dtype=np.int64)
    # a subtree rooted at level L is complete when 8^-(l-L) summed over its leaves reaches 1: exact in units of 8^-maxlevel
    lmax = int(level.max()) if level.size else 0
    if level.size and int(level.min()) < 0:
        raise ValueError("error in levels")
    w = np.left_shift(np.int64(1), 3 * (lmax - level))   # volume of each leaf in finest-cell units
    full = np.int64(1) << (3 * lmax)
    c = np.cumsum(w)
    if level.size == 0 or c[-1] % full:
        raise ValueError("error in levels: the leaves do not fill whole base cells")
    # every base cell must end exactly on a multiple of `full`, and within it the pre-order nesting must hold
    ends = np.nonzero(c % full == 0)[0]
    nbase = int(c[-1] // full)
    if ends.size != nbase:
        raise ValueError("error in levels")
    return nbase


def leaf_centres(nx, level):
    """cell centres in box units [0,1] per leaf, as computeCellCoordinates (hdf42bin.f90:225-269) assigns them"""
    level = np.asarray(level, dtype=np.int64)
    lmax = int(level.max()) if level.size else 0
    w = np.left_shift(np.int64(1), 3 * (lmax - level))
    start = np.cumsum(w) - w                       # Morton-like position of the leaf in finest-cell units
    full = np.int64(1) << (3 * lmax)
    base = start // full
    inner = start % full
    bx, by, bz = base // (nx * nx), (base // nx) % nx, base % nx
    fx = np.zeros(level.size, dtype=np.int64); fy = fx.copy(); fz = fx.copy()
    for b in range(lmax):                          # de-interleave: children are ordered i (x), j (y), k (z)
        oct_ = (inner >> (3 * (lmax - 1 - b))) & 7
        fx = (fx << 1) | (oct_ >> 2); fy = (fy << 1) | ((oct_ >> 1) & 1); fz = (fz << 1) | (oct_ & 1)
    fine = nx << lmax
    half = np.left_shift(np.int64(1), lmax - level)  # leaf edge in finest cells
    x = ((bx << lmax) + fx + 0.5 * half) / fine
    y = ((by << lmax) + fy + 0.5 * half) / fine
    z = ((bz << lmax) + fz + 0.5 * half) / fine
    return x, y, z


def write_cell_array_dat(path, level, x, y, z, HI, HeI, HeII, T, rho):
    n = np.asarray(level).size
    cols = [np.asarray(level, dtype="<i4")] + [np.asarray(a, dtype="<f4") for a in (x, y, z, HI, HeI, HeII, T, rho)]
    if any(c.size != n for c in cols):
        raise ValueError("all nine arrays need one value per leaf")
    write_records(path, cols)


def read_cell_array_dat(path):
    """dict of the nine per-leaf arrays + `nx` (cube root of the number of base cells) + `nleaf`"""
    rec = read_records(path)
    if len(rec) != 9:
        raise ValueError(f"expected 9 records, found {len(rec)}")
    out = {"level": np.frombuffer(rec[0], dtype="<i4").copy()}
    for name, b in zip(CELL_ARRAY_FIELDS[1:], rec[1:]):
        out[name] = np.frombuffer(b, dtype="<f4").copy()
        if out[name].size != out["level"].size:
            raise ValueError(f"record {name} has {out[name].size} values, level has {out['level'].size}")
    nbase = count_base_cells(out["level"])
    nx = int(round(nbase ** (1.0 / 3.0)))
    if nx ** 3 != nbase:
        raise ValueError(f"{nbase} base cells: not a cubic base grid (equiSources.f90:427-436)")
    out["nx"], out["nleaf"] = nx, int(out["level"].size)
    return out


def transport_inputs(cells, box_size, abun2=0.02):
    """keyword arguments of `Transport.set_grid` from a cell array (fp64, leaf order kept); abun2 defaults to the
    value the reference assigns when the grid carries no metals (equiSources.f90:1957-1959)"""
    f64 = lambda a: np.asarray(a, dtype=np.float64)
    n = cells["level"].size
    ab = np.full(n, float(abun2)) if np.isscalar(abun2) else f64(abun2)
    return dict(nx=int(cells["nx"]), level=cells["level"].astype(np.int8), HI=f64(cells["HI"]), HeI=f64(cells["HeI"]),
                HeII=f64(cells["HeII"]), rho=f64(cells["rho"]), abun2=ab, box_size=float(box_size))


def write_grid_dat(path, levels):
    """levels: list of dicts with pos[ncell,3], lT, lnH, lx [, abun[ncell,4]][, vel[ncell,3]] (bin2hdf4.f90:43-83)"""
    rec = [np.array([len(levels)], dtype="<i4")]
    for lv in levels:
        pos = np.asarray(lv["pos"], dtype="<f4")
        n = pos.shape[0]
        rec.append(np.array([n], dtype="<i4"))
        rec += [pos[:, 0], pos[:, 1], pos[:, 2]] + [np.asarray(lv[k], dtype="<f4") for k in ("lT", "lnH", "lx")]
        if "abun" in lv:
            ab = np.asarray(lv["abun"], dtype="<f4")
            rec += [ab[:, i] for i in range(4)]
        if "vel" in lv:
            v = np.asarray(lv["vel"], dtype="<f4")
            rec += [v[:, i] for i in range(3)]
    write_records(path, rec)


def read_grid_dat(path, metals=False, kinematics=False):
    """per-level lists of the input grid; `metals` / `kinematics` as the driver infers them from 'met' / 'vel' in the
    grid name (equiSources.f90:144-158)"""
    rec = read_records(path)
    it = iter(rec)
    nlevels = int(np.frombuffer(next(it), dtype="<i4")[0])
    out = []
    f4 = lambda: np.frombuffer(next(it), dtype="<f4").copy()
    for _ in range(nlevels):
        n = int(np.frombuffer(next(it), dtype="<i4")[0])
        lv = {"pos": np.stack([f4(), f4(), f4()], axis=1), "lT": f4(), "lnH": f4(), "lx": f4()}
        if metals:
            lv["abun"] = np.stack([f4() for _ in range(4)], axis=1)
        if kinematics:
            lv["vel"] = np.stack([f4() for _ in range(3)], axis=1)
        if any(np.asarray(v).shape[0] != n for v in lv.values()):
            raise ValueError("record length differs from ncell")
        out.append(lv)
    if next(it, None) is not None:
        raise ValueError("trailing records: wrong metals / kinematics flags?")
    return out
