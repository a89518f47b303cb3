#!/bin/bash
# 1-GPU job: cells tuning, fixed tests, default bench line (with secondaries), shard / zone timings
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python tools/tune_cells.py > gpurun_out/r02b_tune_cells.log 2>&1
python -m pytest tests/test_diffuse_gpu.py -m gpu -x -q -s 2>&1 | tail -15 > gpurun_out/r02b_pytest_diffuse.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err
python tools/shard_times.py 8 > gpurun_out/r02b_shard_times_8.log 2>&1
python tools/zone_times.py 256 > gpurun_out/r02b_zone_times.log 2>&1
cat gpurun_out/r02b_tune_cells.log; tail -3 gpurun_out/r02b_pytest_diffuse.log; head -c 1500 gpurun_out/r02b_bench_n1.json; tail -3 gpurun_out/r02b_bench_n1.err; tail -3 gpurun_out/r02b_shard_times_8.log
