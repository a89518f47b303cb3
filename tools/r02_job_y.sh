#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_diffuse_amr_gpu.py tests/test_multi_gpu.py -m gpu -x -q --durations=3 > gpurun_out/r02y_pytest_amr.log 2>&1
tail -n 8 gpurun_out/r02y_pytest_amr.log
timeout 900 python tools/bench_amr_unbalanced.py > gpurun_out/r02y_amr_unbalanced.log 2>&1
tail -n 3 gpurun_out/r02y_amr_unbalanced.log
