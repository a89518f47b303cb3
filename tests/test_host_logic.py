"""CPU-only checks: the C-ABI library loads and exports every symbol of include/rtb200.h, the host-side geometry
tables are bit-identical to the oracle's, sharding covers every direction once, and the product fails loudly
without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol(build_product):
    hdr = open(os.path.join(ROOT, "include", "rtb200.h")).read()
    declared = set(re.findall(r"\b(rtb200_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"rtb200_ctx"}
    assert len(declared) >= 15
    L = C.CDLL(build_product)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/rtb200.h but not exported"
    from radiativetransfer_b200 import _lib
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.lib().rtb200_version() == 100


def test_no_cpu_fallback(build_product):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import radiativetransfer_b200 as rt
    with pytest.raises(rt.RTB200Error) as e:
        rt.Transport(device=0)
    assert e.value.status == 13  # RTB200_ERR_CUDA


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "radiativetransfer_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh", ".f90")):
                src = open(os.path.join(dp, f)).read()
                assert "ftte_oracle" not in src and "oracle/" not in src.replace("# oracle/", ""), f


def test_directions_bit_identical_to_oracle(build_product, oracle):
    import radiativetransfer_b200 as rt
    for lvl in (1, 2, 3, 4):
        for r in range(12 * 4 ** (lvl - 1)):
            st, iz, phi, th = oracle.direction(lvl, r)
            assert st == 0
            assert rt.direction(lvl, r) == (iz, phi, th)  # exact equality of the doubles


@pytest.mark.parametrize("n", [7, 32, 128])
def test_patterns_bit_identical_to_oracle(build_product, oracle, uvbg, n):
    import radiativetransfer_b200 as rt
    g = oracle.OracleGrid(1, np.zeros(1, dtype=np.int8), np.ones(1))  # only used for 1-cell traces below
    big = oracle.OracleGrid(n, np.zeros(n ** 3, dtype=np.int8), np.zeros(n ** 3)) if n <= 32 else None
    for r in range(0, 192, 1 if n <= 32 else 7):
        mine = rt.patterns(3, r, n)
        if big is not None:
            o = big.diffuse(uvbg["uvb"], uvbg["beta"], ray_begin=r, ray_end=r + 1, trace_ray=r)
            assert o["status"] == 0
            assert np.array_equal(mine, o["pattern"])
        # invariants at any size
        assert np.allclose(mine[:, 2] + mine[:, 5] + mine[:, 8], 1 / np.sin(rt.direction(3, r)[2]), rtol=1e-12)
    del g


def test_sharding_partitions_all_directions(build_product):
    from radiativetransfer_b200 import sharding
    zone, cost = sharding.direction_costs(3, 16)
    for world in (1, 2, 3, 4, 8):
        shards = sharding.shard_directions(world, nx=16)
        allr = np.concatenate(shards)
        assert np.array_equal(np.sort(allr), np.arange(192))
        loads = np.array([cost[s].sum() for s in shards])
        assert loads.max() / loads.mean() < 1.06, (world, loads)
        again = sharding.shard_directions(world, nx=16)
        assert all(np.array_equal(a, b) for a, b in zip(shards, again))       # a pure function of its arguments
    # a cost factor for the zones sweeping along z moves directions away from the ranks that hold them
    a = sharding.shard_directions(4, nx=16)
    b = sharding.shard_directions(4, nx=16, zone_cost=[1.0, 1.0, 1.5])
    assert np.array_equal(np.sort(np.concatenate(b)), np.arange(192))
    assert any(not np.array_equal(x, y) for x, y in zip(a, b))


def test_workload_leaf_order_matches_preorder(oracle):
    from radiativetransfer_b200 import workloads as W
    g = W.nested_grid(4, 2, W.central_box_refine(0.25, 0.75, levels=2), seed=1)
    # the oracle rebuilds the tree from `level` alone and fails on an inconsistent pre-order
    og = oracle.OracleGrid(4, g["level"], g["HI"], box_size=g["box_size"])
    assert og.nleaf == g["level"].size
    # brute-force pre-order
    lv = []
    def rec(level, x, y, z):
        size = 1.0 / (4 << level)
        c = ((x + .5) * size, (y + .5) * size, (z + .5) * size)
        if level < 2 and all(0.25 <= v < 0.75 for v in c):
            for q in range(8):
                rec(level + 1, 2 * x + (q >> 2), 2 * y + ((q >> 1) & 1), 2 * z + (q & 1))
        else:
            lv.append(level)
    for i in range(4):
        for j in range(4):
            for k in range(4):
                rec(0, i, j, k)
    assert np.array_equal(np.array(lv, dtype=np.int8), g["level"])


def test_point_entry_points_reject_bad_arguments_without_a_device(build_product):
    """argument checks that come before any CUDA call: usable on a CPU-only box"""
    import ctypes as C
    from radiativetransfer_b200 import _lib
    L = _lib.lib()
    n = C.c_int64(0)
    assert L.rtb200_point(None, 0, None, None, None, 0.0, None, 0, 6, 0, None, None, None, None, None, None, None, None,
                          None, None, None, None, None, C.byref(n)) == 12    # RTB200_ERR_ARG
    assert L.rtb200_point_device(None, 0, None, None, None, 0.0, None, 0, 6, 0, None, None, None, None, None, None,
                                 None, None, None, C.byref(n)) == 12
    assert L.rtb200_point_tables(None, 0, None, None, None, 0.0, None, 1, 0.0, None) == 12
    assert b"idepth" in L.rtb200_status_string(11)


def test_every_tuning_key_is_documented_in_the_header():
    """rtb200_set_tuning keys accepted by the library (csrc/api.cu, csrc/multi.cu) = keys listed in include/rtb200.h"""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = "".join(open(os.path.join(root, "radiativetransfer_b200", "csrc", f)).read() for f in ("api.cu", "multi.cu"))
    keys = sorted(set(re.findall(r'k == "([a-z0-9_]+)"', src)))
    assert len(keys) > 20
    hdr = open(os.path.join(root, "include", "rtb200.h")).read()
    missing = [k for k in keys if f'"{k}"' not in hdr and not (k.startswith("zone_cost_") and f'"_{k[-1]}"' in hdr)]
    assert not missing, missing
