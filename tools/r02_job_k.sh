#!/bin/bash
# 1-GPU job: streamed nested-grid sweep: parity tests, timing against the per-wave launches
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_diffuse_amr_gpu.py -m gpu -x -q > gpurun_out/r02k_pytest.log 2>&1
tail -5 gpurun_out/r02k_pytest.log
timeout 600 python tools/bench_amr_variants.py > gpurun_out/r02k_amr_variants.log 2>&1
cat gpurun_out/r02k_amr_variants.log | tail -12
