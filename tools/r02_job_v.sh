#!/bin/bash
# 1-GPU job: the benchmarked 256^3 x 192-direction solve in full against the CPU oracle (all host threads)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nproc > gpurun_out/r02v_slow_test.log
RTB_SLOW_TESTS=1 timeout 2400 python -m pytest tests/test_diffuse_gpu.py -m gpu -x -q -s -k "headline_256cube_full" >> gpurun_out/r02v_slow_test.log 2>&1
tail -n 6 gpurun_out/r02v_slow_test.log
