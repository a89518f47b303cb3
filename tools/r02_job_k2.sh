#!/bin/bash
# 1-GPU job: nested-grid sweep defaults (timing), then one ncu capture of the streamed whole-sweep kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python tools/bench_amr.py 64 3 > gpurun_out/r02k2_bench_amr64.log 2>&1 || exit 1
timeout 600 python tools/bench_amr.py 128 2 > gpurun_out/r02k2_bench_amr128.log 2>&1 || exit 1
tail -2 gpurun_out/r02k2_bench_amr64.log gpurun_out/r02k2_bench_amr128.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:amr_stream --launch-skip 1 -c 1 -f -o gpurun_out/r02k2_amr_stream python tools/bench_amr.py 64 3 > gpurun_out/r02k2_ncu.log 2>&1
echo "ncu rc=$?"
