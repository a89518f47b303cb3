#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python tools/bench_amr_variants.py > gpurun_out/r02w_amr_variants.log 2>&1
cat gpurun_out/r02w_amr_variants.log | tail -12
