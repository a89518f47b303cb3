#!/bin/bash
# 1-GPU job: full GPU test suite, nested-grid timing, default bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02l_pytest_gpu.log 2>&1
tail -4 gpurun_out/r02l_pytest_gpu.log
timeout 600 python tools/bench_amr.py 64 3 > gpurun_out/r02l_bench_amr64.log 2>&1
timeout 600 python tools/bench_amr.py 128 2 > gpurun_out/r02l_bench_amr128.log 2>&1
tail -n 1 gpurun_out/r02l_bench_amr64.log; tail -n 1 gpurun_out/r02l_bench_amr128.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02l_bench_n1.json 2> gpurun_out/r02l_bench_n1.err
echo "bench rc=$?"; head -c 400 gpurun_out/r02l_bench_n1.json; echo
