#!/bin/bash
# 2-GPU job: device-group tests (one process / one process per GPU), N=2 bench lines, plus 1-GPU experiments
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02d_gpus.log 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_formats.py -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/r02d_pytest_multi.log
timeout 300 python tools/bench_amr_variants.py > gpurun_out/r02d_amr_variants.log 2>&1
timeout 300 python -m pytest tests/test_point_gpu.py -m gpu -x -q -k "planned or segmented" 2>&1 | tail -15 > gpurun_out/r02d_pytest_planned.log
timeout 300 python tools/bench_point_modes.py > gpurun_out/r02d_point_modes.log 2>&1
timeout 300 python tools/shard_times.py 8 > gpurun_out/r02d_shard_times_8.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02d_bench_n2.json 2> gpurun_out/r02d_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --workload point-128^3-amr-100src > gpurun_out/r02d_bench_n2_point.json 2> gpurun_out/r02d_bench_n2_point.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --workload combined-64^3-amr3-192dir-64src > gpurun_out/r02d_bench_n2_combined.json 2> gpurun_out/r02d_bench_n2_combined.err
tail -12 gpurun_out/r02d_pytest_multi.log; cat gpurun_out/r02d_amr_variants.log; tail -4 gpurun_out/r02d_pytest_planned.log; cat gpurun_out/r02d_point_modes.log; tail -2 gpurun_out/r02d_shard_times_8.log
for f in gpurun_out/r02d_bench_n2.json gpurun_out/r02d_bench_n2_point.json gpurun_out/r02d_bench_n2_combined.json; do head -c 700 $f; echo; done
tail -5 gpurun_out/r02d_bench_n2.err
