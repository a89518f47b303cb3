#!/bin/bash
# 1-GPU job: full GPU suite, nested-grid / point / shard experiments, default bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02e_pytest_gpu.log
timeout 300 python tools/bench_amr_variants.py > gpurun_out/r02e_amr_variants.log 2>&1
timeout 300 python tools/bench_point_modes.py > gpurun_out/r02e_point_modes.log 2>&1
timeout 400 python tools/shard_times.py 8 0 8 4 2 0:0:24 2:0:24 > gpurun_out/r02e_shard_times_8.log 2>&1
timeout 300 python tools/shard_times.py 1 0 8 4 > gpurun_out/r02e_shard_times_1.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err
tail -6 gpurun_out/r02e_pytest_gpu.log; cat gpurun_out/r02e_amr_variants.log; cat gpurun_out/r02e_point_modes.log; grep "sweep max" gpurun_out/r02e_shard_times_8.log gpurun_out/r02e_shard_times_1.log; head -c 400 gpurun_out/r02e_bench_n1.json; tail -3 gpurun_out/r02e_bench_n1.err
