"""Ionisation equilibrium (solveRateEquations, equiSources.f90:3459-3677): oracle sanity checks on the CPU, and the
CUDA kernel against the oracle -- bit for bit, because the kernel is IEEE +,-,*,/ in the reference's association order
and the only transcendental, log(tgas), is taken on the host with the same libm."""
import numpy as np
import pytest

from radiativetransfer_b200 import workloads as W


def _gas(n, seed, levels=0):
    if levels:
        g = W.nested_grid(n, levels, W.central_box_refine(0.25, 0.75, levels=levels), seed=seed)
    else:
        g = W.uniform_grid(n, seed=seed)
    rng = np.random.default_rng(seed + 1)
    g["tgas"] = 10.0 ** rng.uniform(3.5, 5.5, g["level"].size)
    return g


def test_oracle_equilibrium_properties(oracle):
    """weak radiation: the result does not depend on the starting state, cold gas stays neutral; strong radiation:
    everything ionised; fractions stay in [0, 1]; with no radiation at all cold gas rounds to HI/nH = 1 + ulp and the
    restatement reports what the reference does there (print and stop, equiSources.f90:3634-3640)"""
    kt = W.rate_tables(600)
    n = 6
    g = _gas(n, 3)
    N = n ** 3
    weak = np.full((3, N), 1e-31)
    ksi = np.full(6, 1e6)                                   # photo-rates ~ 4 pi J ksi ~ 1e-24 / s: well below n k2
    args = (n, g["box_size"], g["level"], g["rho"], g["tgas"])
    a = oracle.chemistry(*args, g["HI"], g["HeI"], g["HeII"], kt, J=weak, ksi=ksi)
    b = oracle.chemistry(*args, 0.3 * g["HI"], 0.5 * g["HeI"], g["HeII"], kt, J=weak, ksi=ksi)
    assert a["status"] == 0 and b["status"] == 0
    assert np.array_equal(a["HI"], b["HI"]) and np.array_equal(a["HeI"], b["HeI"])
    nh = W.PSI * g["rho"] / W.MP
    nhe = (1 - W.PSI) * g["rho"] / W.MHE
    assert np.all((a["HI"] >= 0) & (a["HI"] <= nh)) and np.all(a["HeI"] + a["HeII"] <= nhe * (1 + 1e-9))
    cold = g["tgas"] < 8e3
    assert cold.any() and np.all(a["HI"][cold] / nh[cold] > 0.9)
    c = oracle.chemistry(*args, g["HI"], g["HeI"], g["HeII"], kt, J=np.full((3, N), 1e-15), ksi=ksi)
    assert c["status"] == 0 and np.all(c["HI"] / nh < 1e-3) and c["max_change"] > 0
    d = oracle.chemistry(*args, g["HI"], g["HeI"], g["HeII"], kt, J=np.zeros((3, N)), ksi=ksi)
    assert d["status"] in (0, 15)


@pytest.mark.gpu
@pytest.mark.parametrize("n,levels,mode,pscale", [(8, 0, "diffuse", 0.), (6, 2, "diffuse+point", 1e-15),
                                                  (6, 2, "diffuse+point", 1e-13), (10, 0, "uniform", 0.),
                                                  (5, 1, "point", 1e-15)])
def test_gpu_matches_oracle_bit_for_bit(build_product, oracle, uvbg, n, levels, mode, pscale):
    import torch
    import radiativetransfer_b200 as rt
    kt = W.rate_tables(1200)
    g = _gas(n, 11 + n, levels)
    N = g["level"].size
    rng = np.random.default_rng(n)
    t = rt.Transport(device=0)
    t.set_grid(g["nx"], g["level"], g["HI"], g["HeI"], g["HeII"], g["rho"], g["abun2"], g["box_size"])
    t.set_rate_tables(kt)
    t.set_temperature(g["tgas"])
    J = rates = None
    ksi = np.concatenate([uvbg["ksi24"], uvbg["ksi25"], uvbg["ksi26"]])
    uniform = None
    if "diffuse" in mode:
        J = uvbg["uvb"][:, None] * rng.uniform(0.0, 1.0, (3, N))
    if "point" in mode:
        cell = g["box_size"] / (g["nx"] * 2.0 ** g["level"])
        rates = np.zeros((6, N))
        rates[:3] = rng.uniform(0, pscale, (3, N)) * cell ** 3 * np.stack([g["HI"], np.maximum(g["HeII"], 1e-30), g["HeI"]])
    if J is None:
        uniform = np.array([3e-13, 1e-15, 2e-13, 0.2 * g["box_size"] / g["nx"]])
    Jd = torch.from_numpy(J).cuda() if J is not None else None
    Rd = torch.from_numpy(rates).cuda() if rates is not None else None
    o = oracle.chemistry(g["nx"], g["box_size"], g["level"], g["rho"], g["tgas"], g["HI"], g["HeI"], g["HeII"], kt,
                         rates=rates, J=J, ksi=ksi, uniform=uniform)
    if o["status"] != 0:
        # the reference's bisection leaves [0, 1] for this input and the driver stops (equiSources.f90:3634-3655):
        # the GPU path reports the same condition
        with pytest.raises(rt.RTB200Error) as e:
            t.chemistry_device(Rd.data_ptr() if Rd is not None else 0, Jd.data_ptr() if Jd is not None else 0,
                               ksi=ksi, uniform=uniform)
        assert e.value.status == o["status"] == 15
        t.close()
        return
    change = t.chemistry_device(Rd.data_ptr() if Rd is not None else 0, Jd.data_ptr() if Jd is not None else 0,
                                ksi=ksi, uniform=uniform)
    HI, HeI, HeII = t.get_species()
    t.close()
    assert np.array_equal(HI, o["HI"]) and np.array_equal(HeI, o["HeI"]) and np.array_equal(HeII, o["HeII"])
    assert change == o["max_change"]


@pytest.mark.gpu
def test_outer_iteration_stays_on_the_device(build_product, oracle, uvbg):
    """three passes of the reference's outer loop (diffuse sweep -> chemistry) with J, HI, HeI, HeII resident on the
    GPU, against the same loop on the CPU oracle"""
    import torch
    import radiativetransfer_b200 as rt
    kt = W.rate_tables(1200)
    n = 8
    g = _gas(n, 5)
    ksi = np.concatenate([uvbg["ksi24"], uvbg["ksi25"], uvbg["ksi26"]])
    t = rt.Transport(device=0, math=rt.MATH_FAITHFUL)
    t.set_grid(g["nx"], g["level"], g["HI"], g["HeI"], g["HeII"], g["rho"], g["abun2"], g["box_size"])
    t.set_rate_tables(kt); t.set_temperature(g["tgas"])
    Jd = torch.zeros(3, n ** 3, dtype=torch.float64, device="cuda:0")
    s = torch.cuda.current_stream().cuda_stream
    og = oracle.OracleGrid(n, g["level"], g["HI"], g["HeI"], g["HeII"], box_size=g["box_size"])
    HI, HeI, HeII = g["HI"].copy(), g["HeI"].copy(), g["HeII"].copy()
    for it in range(3):
        t.diffuse_device(uvbg["uvb"], uvbg["beta"], Jd.data_ptr(), stream=s)
        change = t.chemistry_device(0, Jd.data_ptr(), ksi=ksi, stream=s)
        o = og.diffuse(uvbg["uvb"], uvbg["beta"])
        c = oracle.chemistry(n, g["box_size"], g["level"], g["rho"], g["tgas"], HI, HeI, HeII, kt, J=o["J"], ksi=ksi)
        assert o["status"] == 0 and c["status"] == 0
        HI, HeI, HeII = c["HI"], c["HeI"], c["HeII"]
        og.set_species(HI, HeI, HeII)
        assert abs(change - c["max_change"]) <= 1e-9 * max(c["max_change"], 1e-30)
    gHI, gHeI, gHeII = t.get_species()
    t.close()
    nh = W.PSI * g["rho"] / W.MP
    assert np.max(np.abs(gHI - HI) / nh) < 1e-9           # J agrees to ~1e-13; the bisection is Lipschitz in the rates
    assert np.max(np.abs(gHeI - HeI) / nh) < 1e-9


def test_oracle_compute_mass_known_answer(oracle):
    """computeMass (equiSources.f90:4369-4393): constant density in a nested grid -> mass = density x box volume,
    whatever the refinement (the leaves tile the box); the neutral mass follows HI the same way"""
    g = _gas(6, 21, levels=2)
    N = g["level"].size
    rho0, hi0 = 3.0e-25, 2.0e-3
    neutral, total = oracle.compute_mass(6, g["box_size"], g["level"], np.full(N, hi0), np.full(N, rho0))
    msun, mh, psi = float(np.float32(1.98892e33)), float(np.float32(1.6726231e-24)), float(np.float32(0.76))
    vol = g["box_size"] ** 3
    assert abs(total - psi * rho0 * vol / msun) <= 1e-12 * total
    assert abs(neutral - hi0 * mh * vol / msun) <= 1e-12 * neutral
    # and a direct numpy evaluation with random fields (pairwise summation: agreement to rounding)
    size = g["box_size"] / (2.0 ** g["level"].astype(np.float64) * 6)
    neutral, total = oracle.compute_mass(6, g["box_size"], g["level"], g["HI"], g["rho"])
    assert abs(total - np.sum(psi * g["rho"] * size ** 3 / msun)) <= 1e-12 * total
    assert abs(neutral - np.sum(g["HI"] * mh * size ** 3 / msun)) <= 1e-12 * neutral


@pytest.mark.gpu
@pytest.mark.parametrize("n,levels", [(8, 0), (6, 2), (40, 0), (1, 0)])
def test_gpu_compute_mass_matches_oracle(build_product, oracle, n, levels):
    """rtb200_compute_mass: same per-leaf terms as the reference, summed as a fixed tree instead of serially ->
    equal to rounding of the summation order, and reproducible; it sees the species the chemistry step left"""
    import radiativetransfer_b200 as rt
    g = _gas(n, 31 + n, levels)
    t = rt.Transport(device=0)
    t.set_grid(g["nx"], g["level"], g["HI"], g["HeI"], g["HeII"], g["rho"], g["abun2"], g["box_size"])
    a = t.compute_mass()
    assert a == t.compute_mass()
    o = oracle.compute_mass(g["nx"], g["box_size"], g["level"], g["HI"], g["rho"])
    assert abs(a[0] - o[0]) <= 1e-12 * o[0] and abs(a[1] - o[1]) <= 1e-12 * o[1]
    t.update_species(HI=0.5 * g["HI"])
    b = t.compute_mass()
    o2 = oracle.compute_mass(g["nx"], g["box_size"], g["level"], 0.5 * g["HI"], g["rho"])
    assert abs(b[0] - o2[0]) <= 1e-12 * o2[0] and b[1] == a[1]
    t.close()
    fresh = rt.Transport(device=0)
    with pytest.raises(rt.RTB200Error):
        fresh.compute_mass()                                  # no grid yet
    fresh.close()
