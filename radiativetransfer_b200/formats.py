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
    _check_volume_range(lmax, level.size)
    _check_preorder(level)
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


def _check_volume_range(lmax, nleaf):
    """leaf volumes are counted in int64 units of 8^-lmax and summed over all leaves: 3*lmax + log2(nleaf) must stay
    below 63 bits (e.g. 128^3 base cells allow 13 levels; the reference tools size for levels 0..15)"""
    if 3 * lmax + max(int(nleaf), 1).bit_length() >= 63:
        raise ValueError(f"levels up to {lmax} with {nleaf} leaves overflow the 64-bit volume count")


def _check_preorder(level):
    """pre-order nesting of a leaf-level sequence (the walk of readCellArray.f90:154-187): after a leaf of level l the
    walk stands at a sibling slot of some level <= l; the next leaf can open deeper levels only by refining that slot"""
    lv = np.asarray(level, dtype=np.int64)
    if lv.size == 0:
        return
    # remaining[k] = children still missing at level k of the current path (k >= 1); a vectorised check is not possible
    # for the nesting itself, but the necessary condition below is cheap and the exact walk is only run when it holds
    remaining = [0] * (int(lv.max()) + 2)
    depth = 0
    for l in lv.tolist():
        if l < depth and any(remaining[k] for k in range(l + 1, depth + 1)):
            raise ValueError("error in levels: a refined cell is closed before its 8 children are complete")
        while depth < l:            # refine the current slot down to level l
            depth += 1
            remaining[depth] = 8
        if l > 0:
            remaining[l] -= 1
        depth = l
        while depth > 0 and remaining[depth] == 0:   # a completed octet closes its parent slot
            depth -= 1
            if depth > 0:
                remaining[depth] -= 1
    if depth != 0:
        raise ValueError("error in levels: the last refined cell is incomplete")


def leaf_centres(nx, level):
    """cell centres in box units [0,1] per leaf, as computeCellCoordinates (hdf42bin.f90:225-269) assigns them"""
    level = np.asarray(level, dtype=np.int64)
    lmax = int(level.max()) if level.size else 0
    _check_volume_range(lmax, level.size)
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


# ----------------------------------------------------------------------------------------------------------------------
# octree from the per-level cell lists of the input grid (equiSources.f90:425-618 + placeCellProjectWithVelocity :1870-1974)
# ----------------------------------------------------------------------------------------------------------------------
_F = np.float32
_PSI = float(_F(0.76))                       # definitionsModule.f90:261
_MP = float(_F(1.6726231e-24))
_MN = float(_F(1.67492728e-24))
_MHE = 2.0 * (_MP + _MN)
_KPC = float(_F(1.e3)) * float(_F(3.08568025e18))
_STATE = ("tgas", "rho", "HI", "HeI", "HeII", "abun2")


def _pow10(a):
    """10.**x through the C library's pow (what gfortran calls), element by element: numpy's vectorised power loop can
    differ from it in the last bit"""
    a = np.asarray(a, dtype=np.float32).astype(np.float64)
    return np.frompyfunc(lambda v: 10.0 ** v, 1, 1)(a).astype(np.float64)


def smooth_level1_abundance(abun2, idx, nx):
    """Level-1 smoothing of the second abundance (equiSources.f90:527-578): the values are scattered onto the n^3 base
    grid (a later duplicate overwrites), filtered twice with (1/4, 1/2, 1/4) along i, then j, then k -- nothing comes
    in across the box faces -- and gathered back into the real*4 list.  Each output adds its three terms in the order
    the reference's scatter loop produces them: 0.25 u(i-1), then 0.5 u(i), then 0.25 u(i+1), starting from 0."""
    u = np.zeros((nx, nx, nx))
    u[idx[:, 0], idx[:, 1], idx[:, 2]] = np.asarray(abun2, dtype=np.float32).astype(np.float64)
    for _ in range(2):
        for ax in range(3):
            t = np.zeros_like(u)
            lo = [slice(None)] * 3; hi = [slice(None)] * 3
            lo[ax] = slice(0, nx - 1); hi[ax] = slice(1, nx)
            t[tuple(hi)] = t[tuple(hi)] + 0.25 * u[tuple(lo)]      # from i-1
            t = t + 0.5 * u                                        # own
            t[tuple(lo)] = t[tuple(lo)] + 0.25 * u[tuple(hi)]      # from i+1
            u = t
    return u[idx[:, 0], idx[:, 1], idx[:, 2]].astype(np.float32)


def build_leaves(levels, metals=False):
    """Leaf arrays in `writeCell` order from the per-level lists of `read_grid_dat` (positions in kpc).

    Follows the driver: level 1 must hold n^3 cells (:427-436); the box is the level-1 min/max stretched by n/(n-1)
    (:455-477) and positions are normalised and stored back in SINGLE precision (:483-489); a level-l cell is dropped
    into the tree by descending l-1 times with `.lt.0.5` tests, and children created on the way inherit tgas, rho, HI,
    HeI, HeII of their parent and get abun2 = 0 (:1882-1932); with metals the level-1 abun(:,2) is smoothed on the
    base grid first (:527-578, `smooth_level1_abundance`); the target cell takes tgas = 10**lT, nH = 10**lnH,
    HI = nH 10**lx, rho = nH mh/psi, HeI = (1-psi) rho/mhe, HeII = 0, abun2 = abun(:,2) or 0.02 (:1935-1959).
    Returns the keyword arguments of `Transport.set_grid` plus `tgas`."""
    n1 = levels[0]["pos"].shape[0]
    nx = int(round(n1 ** (1.0 / 3.0)))
    if nx ** 3 != n1:
        raise ValueError("base grid needs to be of size n^3")
    p1 = np.asarray(levels[0]["pos"], dtype=np.float32).astype(np.float64)
    lo, hi = p1.min(axis=0), p1.max(axis=0)
    mid, half = 0.5 * (lo + hi), 0.5 * (hi - lo) * float(nx) / float(nx - 1) if nx > 1 else 0.5 * (hi - lo)
    a, b = mid - half, mid + half
    box_size = abs(a[0] - b[0]) * _KPC
    # per depth: packed integer coordinates -> row of the depth's state table
    index = [dict() for _ in levels]
    state = [dict((k, np.zeros(0)) for k in _STATE) for _ in levels]
    refined = [set() for _ in levels]            # packed coordinates of refined nodes per depth

    def pack(ix, iy, iz, depth):
        m = nx << depth
        return (ix * m + iy) * m + iz

    def append(depth, keys, vals):
        base = state[depth]["rho"].size
        for j, k in enumerate(keys):
            index[depth][int(k)] = base + j
        for name in _STATE:
            state[depth][name] = np.concatenate([state[depth][name], vals[name]])

    # depth 0: every base cell exists, zero state (:495-520)
    ii, jj, kk = np.meshgrid(np.arange(nx), np.arange(nx), np.arange(nx), indexing="ij")
    append(0, pack(ii.ravel(), jj.ravel(), kk.ravel(), 0), dict((k, np.zeros(nx ** 3)) for k in _STATE))
    for depth, lv in enumerate(levels):
        pos = np.asarray(lv["pos"], dtype=np.float32).astype(np.float64)
        pn = ((pos - a) / (b - a)).astype(np.float32).astype(np.float64)      # stored back as real*4
        base_i = (pn * nx).astype(np.int64)                                    # int(x0*nx)
        if np.any(base_i < 0) or np.any(base_i >= nx):
            raise ValueError("cell outside the box")
        frac = pn * float(nx) - base_i.astype(np.float64)
        coords = [base_i.copy()]
        for _ in range(depth):                                                 # `.lt.0.5` descents
            bit = (frac >= 0.5).astype(np.int64)
            frac = 2.0 * frac - bit
            coords.append(2 * coords[-1] + bit)
        # create the path: the ancestor at depth t-1 must be refined, missing children inherit its state
        for t in range(1, depth + 1):
            par = np.unique(pack(coords[t - 1][:, 0], coords[t - 1][:, 1], coords[t - 1][:, 2], t - 1))
            new = np.array([p for p in par if int(p) not in refined[t - 1]], dtype=np.int64)
            if new.size == 0:
                continue
            for p in new:
                refined[t - 1].add(int(p))
            m = nx << (t - 1)
            px, py, pz = new // (m * m), (new // m) % m, new % m
            rows = np.array([index[t - 1][int(p)] for p in new])
            o = np.arange(8)
            cx = (2 * px[:, None] + (o >> 2)).ravel(); cy = (2 * py[:, None] + ((o >> 1) & 1)).ravel()
            cz = (2 * pz[:, None] + (o & 1)).ravel()
            vals = dict((k, np.repeat(state[t - 1][k][rows], 8)) for k in _STATE)
            vals["abun2"] = np.zeros(8 * new.size)                             # :1907 children start with abun2 = 0
            append(t, pack(cx, cy, cz, t), vals)
        # assign the cells of this level (a later duplicate overwrites an earlier one, as the sequential loop does)
        abun = np.asarray(lv["abun"], dtype=np.float32) if metals else None
        if metals and depth == 0:
            abun = abun.copy()
            abun[:, 1] = smooth_level1_abundance(abun[:, 1], base_i, nx)
        nH = _pow10(lv["lnH"])
        rho = nH * _MP / _PSI
        vals = dict(tgas=_pow10(lv["lT"]), rho=rho, HI=nH * _pow10(lv["lx"]),
                    HeI=(1.0 - _PSI) * rho / _MHE * 1.0, HeII=np.zeros(nH.size),
                    abun2=abun[:, 1].astype(np.float64) if metals
                    else np.full(nH.size, float(_F(0.02))))
        keys = pack(coords[depth][:, 0], coords[depth][:, 1], coords[depth][:, 2], depth)
        rows = np.array([index[depth][int(k)] for k in keys])
        for name in _STATE:
            state[depth][name][rows] = vals[name]
    # leaves in pre-order: sort by (base cell, octant digits)
    lmax = len(levels) - 1
    out_keys, out_level, out_vals = [], [], dict((k, []) for k in _STATE)
    for depth in range(len(levels)):
        if not index[depth]:
            continue
        keys = np.array([k for k in index[depth] if k not in refined[depth]], dtype=np.int64)
        if keys.size == 0:
            continue
        rows = np.array([index[depth][int(k)] for k in keys])
        m = nx << depth
        x, y, z = keys // (m * m), (keys // m) % m, keys % m
        key = ((x >> depth) * nx + (y >> depth)) * nx + (z >> depth)
        for d in range(depth - 1, -1, -1):
            key = (key << 3) | ((((x >> d) & 1) << 2) | (((y >> d) & 1) << 1) | ((z >> d) & 1))
        out_keys.append(key << (3 * (lmax - depth)))
        out_level.append(np.full(keys.size, depth, dtype=np.int8))
        for name in _STATE:
            out_vals[name].append(state[depth][name][rows])
    order = np.argsort(np.concatenate(out_keys), kind="stable")
    res = dict((k, np.concatenate(v)[order]) for k, v in out_vals.items())
    return dict(nx=nx, level=np.concatenate(out_level)[order], HI=res["HI"], HeI=res["HeI"], HeII=res["HeII"],
                rho=res["rho"], abun2=res["abun2"], box_size=box_size, tgas=res["tgas"])


# ------------------------------------------------------------------------------------------------------
# flat dataset container of the gfortran recipe (the HDF4 stand-in hdf4_stub.c under ref_harness, test infrastructure) (stands in for HDF4 SD files where the unmodified reference
# driver is built without libmfhdf): "RTBSD001" | int32 nsds | nsds x { name[64] | type | rank | dims[4] | nbytes | data }
# ------------------------------------------------------------------------------------------------------
_SD_TYPES = {np.dtype("<i4"): 24, np.dtype("<f4"): 5, np.dtype("<f8"): 6}
_SD_DTYPES = {24: "<i4", 25: "<u4", 5: "<f4", 6: "<f8", 20: "i1", 21: "u1"}


def write_sd_container(path, datasets):
    """datasets = list of (name, array); arrays are stored in Fortran (column-major) order with dims as Fortran sees them"""
    with open(path, "wb") as f:
        f.write(b"RTBSD001")
        f.write(np.array([len(datasets)], dtype="<i4").tobytes())
        for name, a in datasets:
            a = np.asarray(a)
            if a.dtype not in _SD_TYPES:
                raise ValueError(f"dataset {name}: unsupported dtype {a.dtype}")
            if a.ndim < 1 or a.ndim > 4:
                raise ValueError("rank 1..4 expected")
            dims = list(a.shape) + [0] * (4 - a.ndim)
            data = np.asfortranarray(a).tobytes(order="F")
            f.write(name.encode()[:63].ljust(64, b"\0"))
            f.write(np.array([_SD_TYPES[a.dtype], a.ndim] + dims, dtype="<i4").tobytes())
            f.write(np.array([len(data)], dtype="<i8").tobytes())
            f.write(data)


def read_sd_container(path):
    """list of (name, array) in file order"""
    out = []
    with open(path, "rb") as f:
        if f.read(8) != b"RTBSD001":
            raise ValueError("not a dataset container")
        n = int(np.frombuffer(f.read(4), dtype="<i4")[0])
        for _ in range(n):
            name = f.read(64).split(b"\0", 1)[0].decode()
            hdr = np.frombuffer(f.read(24), dtype="<i4")
            nbytes = int(np.frombuffer(f.read(8), dtype="<i8")[0])
            typ, rank = int(hdr[0]), int(hdr[1])
            dims = [int(x) for x in hdr[2:2 + rank]]
            a = np.frombuffer(f.read(nbytes), dtype=_SD_DTYPES[typ])
            if int(np.prod(dims)) == a.size:
                a = a.reshape(dims, order="F")
            out.append((name, a.copy()))
    return out


def write_grid_container(path, levels, metals=True, kinematics=False):
    """the reference's input grid `<grid>.h4` (equiSources.f90:316-423; written by bin2hdf4.f90:106-165) as a container:
    dataset 0 `nlevels`, then per level pos[ncell,3], lT, lnH, lx (float32) [, abun[ncell,4]] [, vel[ncell,3]].
    levels = list of dicts with those keys (level 1 first, exactly n^3 cells)."""
    ds = [("nlevels", np.array([len(levels)], dtype="<i4"))]
    for i, lv in enumerate(levels, 1):
        ds.append((f"pos{i}", np.asarray(lv["pos"], dtype="<f4")))
        for k in ("lT", "lnH", "lx"):
            ds.append((f"{k}{i}", np.asarray(lv[k], dtype="<f4")))
        if metals:
            ds.append((f"abun{i}", np.asarray(lv["abun"], dtype="<f4")))
        if kinematics:
            ds.append((f"vel{i}", np.asarray(lv["vel"], dtype="<f4")))
    write_sd_container(path, ds)


def build_leaves_device(levels, metals=False, device=0, _library=None, _prefix="rtb200_octree_"):
    """`build_leaves` on the GPU: rtb200_octree_build (csrc/octree_build.cu) -- the same octree from the same per-level
    lists, built with sorted key sets instead of cell-by-cell insertion.  Returns the keyword arguments of
    `Transport.set_grid` plus `tgas`.  (`_library` / `_prefix`: the test suite's host-backend build of the same source.)"""
    import ctypes as C
    if _library is None:
        from . import _lib
        L = _lib.lib()
    else:
        L = _library
    nl = len(levels)
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    pos = [f32(lv["pos"]) for lv in levels]
    lT = [f32(lv["lT"]) for lv in levels]
    lnH = [f32(lv["lnH"]) for lv in levels]
    lx = [f32(lv["lx"]) for lv in levels]
    ab = [f32(np.asarray(lv["abun"])[:, 1]) for lv in levels] if metals else None
    ncell = np.array([p.shape[0] for p in pos], dtype=np.int64)
    arr = lambda lst: (C.c_void_p * nl)(*[a.ctypes.data_as(C.c_void_p) for a in lst])
    nleaf, nx, box, h = C.c_int64(0), C.c_int32(0), C.c_double(0), C.c_void_p()
    build = getattr(L, _prefix + "build"); get = getattr(L, _prefix + "get"); free = getattr(L, _prefix + "free")
    build.restype = get.restype = free.restype = C.c_int
    build.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                      C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_void_p)]
    get.argtypes = [C.c_void_p] * 8
    free.argtypes = [C.c_void_p]
    st = build(int(device), nl, ncell.ctypes.data_as(C.c_void_p), arr(pos), arr(lT), arr(lnH), arr(lx),
               arr(ab) if metals else None, C.byref(nleaf), C.byref(nx), C.byref(box), C.byref(h))
    if st:
        if _library is None:
            from . import _lib
            _lib.check(st, "rtb200_octree_build")
        raise RuntimeError(f"octree build failed: status {st}")
    N = nleaf.value
    out = dict(nx=nx.value, box_size=box.value, level=np.zeros(N, dtype=np.int8))
    for k in ("HI", "HeI", "HeII", "rho", "abun2", "tgas"):
        out[k] = np.zeros(N)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    st = get(h, p(out["level"]), p(out["HI"]), p(out["HeI"]), p(out["HeII"]), p(out["rho"]), p(out["abun2"]), p(out["tgas"]))
    free(h)
    if st:
        raise RuntimeError(f"octree get failed: status {st}")
    return out
