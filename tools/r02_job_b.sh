#!/bin/bash
# 1-GPU job: full GPU suite after the nested-grid restructure, nested-grid timings, default bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02c_pytest_gpu.log
python tools/bench_amr.py 64 3 > gpurun_out/r02c_bench_amr.log 2>&1
python tools/bench_amr.py 128 2 >> gpurun_out/r02c_bench_amr.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err
tail -8 gpurun_out/r02c_pytest_gpu.log; cat gpurun_out/r02c_bench_amr.log; head -c 600 gpurun_out/r02c_bench_n1.json; tail -3 gpurun_out/r02c_bench_n1.err
