#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02z_pytest_gpu.log 2>&1
tail -n 4 gpurun_out/r02z_pytest_gpu.log
