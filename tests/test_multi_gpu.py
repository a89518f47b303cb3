"""Device groups (csrc/multi.cu): one C-ABI handle for several GPUs.

A group of ONE device runs on any GPU box and exercises the whole group code path (slab bookkeeping, exchange buffers,
the peer-memory reduce kernel with its fused epilogues, slab chemistry).  The tests marked `needs2` need two devices
(gpurun --gpus 2): one process driving both (rtb200_create_multi) and one process per GPU (rtb200_create_rank under
torch.distributed.run, tests/multi_rank_worker.py).  Bar: the N-device result equals the 1-device result to 1e-13
relative (only the order in which the directions' / sources' contributions are added differs)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, rel_err
from radiativetransfer_b200 import workloads as W

pytestmark = pytest.mark.gpu
S24 = float(np.float32(6.3e-18))


def _ndev():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_ndev() < 2, reason="needs two CUDA devices")


@pytest.fixture(scope="module")
def rt(build_product):
    import radiativetransfer_b200 as rt
    return rt


def _set(t, g):
    t.set_grid(g["nx"], g["level"], g["HI"], g["HeI"], g["HeII"], g["rho"], g["abun2"], g["box_size"])


def _grids():
    return {"uniform-24": W.uniform_grid(24, seed=3),                    # 13824 leaves: divisible by 2, 4, 8
            "nested-odd": W.nested_grid(6, 2, W.central_box_refine(0.2, 0.7, levels=2), seed=5)}   # leaf count not divisible


def _check_group(rt, devices, uvbg, reduce_modes):
    spectra = W.synthetic_spectra()
    ksi = np.concatenate([uvbg["ksi24"], uvbg["ksi25"], uvbg["ksi26"]])
    for name, g in _grids().items():
        N = g["level"].size
        one = rt.Transport(device=devices[0])
        _set(one, g)
        J1, nseg1 = one.diffuse(uvbg["uvb"], uvbg["beta"])
        src = np.array([N // 2, N // 3, 5, N - 7, N // 2 + 11], dtype=np.int32)
        wt = np.array([1, 2, 0, 1, 3], dtype=np.int32)
        one.set_math(rt.MATH_FAITHFUL)
        p1 = one.point(spectra, src, wt, dust_approximation=1, rates=np.full((6, N), 1e-30))
        one.set_math(rt.MATH_FAST)
        for mode in reduce_modes:
            grp = rt.Transport(devices=devices)
            grp.set_tuning(multi_reduce=mode)
            _set(grp, g)
            info = grp.info()
            assert info["nranks"] == len(devices) and info["slab"] * len(devices) >= N
            if len(devices) > 1:
                assert info["reduce_mode"] == mode, "peer access between the devices of one box expected"
            # shards partition the directions
            allr = np.concatenate([grp.shard(r) for r in range(len(devices))])
            assert np.array_equal(np.sort(allr), np.arange(192))
            # host-buffer call: same entry point, same arrays
            J, nseg = grp.diffuse(uvbg["uvb"], uvbg["beta"])
            assert nseg == nseg1
            assert rel_err(J, J1) < 1e-13, (name, mode)
            # new species through the slab upload + all-gather, then again
            grp.update_species(HI=g["HI"] * 0.5); one.update_species(HI=g["HI"] * 0.5)
            Jb, _ = grp.diffuse(uvbg["uvb"], uvbg["beta"])
            Jb1, _ = one.diffuse(uvbg["uvb"], uvbg["beta"])
            assert rel_err(Jb, Jb1) < 1e-13
            HIg = grp.get_species()[0]
            assert np.array_equal(HIg, g["HI"] * 0.5)
            grp.update_species(HI=g["HI"]); one.update_species(HI=g["HI"])
            # resident step: J slab + fused photo-rates
            nsegr = grp.diffuse_resident(uvbg["uvb"], uvbg["beta"], ksi=ksi)
            grp.sync()
            assert nsegr == nseg1
            fourpi = 4.0 * float(np.float32(3.141592654))
            for loc in range(info["nlocal"]):
                off, cnt, Js, Ks, _ = grp.slab_get(loc)
                assert rel_err(Js, J1[:, off:off + cnt]) < 1e-13
                k24 = fourpi * Js[0] * ksi[0] + fourpi * Js[1] * ksi[1] + fourpi * Js[2] * ksi[2]
                k25 = fourpi * Js[2] * ksi[3]
                k26 = fourpi * Js[1] * ksi[4] + fourpi * Js[2] * ksi[5]
                assert rel_err(Ks, np.stack([k24, k25, k26])) < 1e-14
            # run-to-run reproducibility of the peer-memory reduction (fixed summation order)
            if mode == 1:
                J2, _ = grp.diffuse(uvbg["uvb"], uvbg["beta"])
                assert np.array_equal(J, J2)
            # point sources: sharded round robin, accumulated into the caller's rate fields, diagnostics per source
            grp.set_math(rt.MATH_FAITHFUL)
            p = grp.point(spectra, src, wt, dust_approximation=1, rates=np.full((6, N), 1e-30))
            grp.set_math(rt.MATH_FAST)
            assert p["nseg"] == p1["nseg"]
            assert rel_err(p["rates"], p1["rates"], floor=1e-300) < 1e-12
            assert rel_err(p["ndot_remaining"], p1["ndot_remaining"], floor=1e-300) < 1e-12
            assert np.array_equal(p["highest_pixel_level"], p1["highest_pixel_level"])
            grp.close()
        one.close()


def test_group_of_one_device_equals_plain_context(rt, uvbg):
    _check_group(rt, [0], uvbg, reduce_modes=[1, 0])


@needs2
def test_two_devices_one_process(rt, uvbg):
    _check_group(rt, [0, 1], uvbg, reduce_modes=[1, 0])


def _outer_iteration_reference(rt, g, uvbg, ktab, tgas, spectra, src, wt, passes):
    """single device: (point pass ->) sweep -> solveRateEquations, `passes` times"""
    import torch
    N = g["level"].size
    ksi = np.concatenate([uvbg["ksi24"], uvbg["ksi25"], uvbg["ksi26"]])
    one = rt.Transport(device=0)
    _set(one, g)
    one.set_rate_tables(ktab); one.set_temperature(tgas)
    J = torch.zeros(3, N, dtype=torch.float64, device="cuda:0")
    R = torch.zeros(6, N, dtype=torch.float64, device="cuda:0")
    s = torch.cuda.current_stream(0).cuda_stream
    for _ in range(passes):
        R.zero_()
        if src is not None:
            one.point_device(spectra, src, wt, R.data_ptr(), stream=s)
        one.diffuse_device(uvbg["uvb"], uvbg["beta"], J.data_ptr(), stream=s)
        one.chemistry_device(R.data_ptr() if src is not None else 0, J.data_ptr(), ksi=ksi, stream=s, want_change=False)
    torch.cuda.synchronize()
    out = one.get_species()
    one.close()
    return out


@pytest.mark.parametrize("ndev", [1, 2])
def test_outer_iteration_on_slabs(rt, uvbg, ndev):
    """combined step of a device group: point pass + sweep + ionisation equilibrium on each rank's slab + all-gather of
    the new species, three passes, against the single-device loop"""
    if _ndev() < ndev:
        pytest.skip("needs more devices")
    g = W.nested_grid(6, 2, W.central_box_refine(0.2, 0.7, levels=2), seed=5, tau_lo=1e-2, tau_hi=2.0, beta24=S24)
    N = g["level"].size
    ktab = W.rate_tables(500)
    tgas = 10.0 ** np.random.default_rng(2).uniform(4.0, 4.1, N)
    spectra = W.synthetic_spectra()
    src = np.array([N // 2, 17, N - 3], dtype=np.int32); wt = np.array([2, 1, 1], dtype=np.int32)
    ksi = np.concatenate([uvbg["ksi24"], uvbg["ksi25"], uvbg["ksi26"]])
    for with_sources in (False, True):
        ref = _outer_iteration_reference(rt, g, uvbg, ktab, tgas, spectra, src if with_sources else None, wt, 3)
        grp = rt.Transport(devices=list(range(ndev)))
        _set(grp, g)
        grp.set_rate_tables(ktab); grp.set_temperature(tgas)
        for _ in range(3):
            if with_sources:
                grp.point_resident(spectra, src, wt)
            grp.diffuse_resident(uvbg["uvb"], uvbg["beta"], ksi=ksi, chemistry=True)
        grp.sync()
        got = grp.get_species()
        for a, b in zip(got, ref):
            assert rel_err(a, b, floor=1e-300) < 1e-9      # J / rates differ by summation order; equilibrium is smooth
        m_n, m_t = grp.compute_mass()
        assert m_t > 0 and 0 < m_n <= m_t
        grp.close()


@needs2
def test_one_process_per_gpu(rt):
    """rtb200_create_rank under torch.distributed.run, 2 ranks: tests/multi_rank_worker.py compares every rank's slab
    with the single-device result it computes itself"""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(ROOT, "tests", "multi_rank_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    sys.stdout.write(r.stdout[-3000:]); sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0
    assert r.stdout.count("RANK-OK") == 2
