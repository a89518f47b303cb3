#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_diffuse_amr_gpu.py -m gpu -x -q > gpurun_out/r02y_pytest_amr.log 2>&1
tail -n 5 gpurun_out/r02y_pytest_amr.log
timeout 900 python tools/bench_amr_unbalanced.py > gpurun_out/r02y_amr_unbalanced.log 2>&1
tail -n 4 gpurun_out/r02y_amr_unbalanced.log
