#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_diffuse_gpu.py tests/test_chemistry.py -m gpu -x -q > gpurun_out/r02y_pytest_uniform.log 2>&1
tail -n 4 gpurun_out/r02y_pytest_uniform.log
timeout 300 python tools/sanitize_small.py > gpurun_out/r02y_small_calls.log 2>&1; echo "small calls rc=$?"; tail -n 3 gpurun_out/r02y_small_calls.log
