#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_diffuse_amr_gpu.py -m gpu -x -q --durations=5 > gpurun_out/r02y_pytest_amr.log 2>&1
tail -n 12 gpurun_out/r02y_pytest_amr.log
