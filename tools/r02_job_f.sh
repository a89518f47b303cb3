#!/bin/bash
# 8-GPU job: device group at 8 ranks (one process per GPU): worker test, bench lines at N = 8 and 4
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02g_gpus.log 2>&1
nvidia-smi topo -m >> gpurun_out/r02g_gpus.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tests/multi_rank_worker.py > gpurun_out/r02g_worker8.log 2>&1
echo "worker rc=$?" >> gpurun_out/r02g_worker8.log
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02g_bench_n$n.json 2> gpurun_out/r02g_bench_n$n.err
done
RTB200_MULTI_REDUCE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02g_bench_n8_nccl.json 2> gpurun_out/r02g_bench_n8_nccl.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 3 --warmup 3 --workload point-256^3-amr-1000src > gpurun_out/r02g_bench_n8_point1000.json 2> gpurun_out/r02g_bench_n8_point1000.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 8 --steps 5 --warmup 3 --workload iterate10-256^3-uniform-192dir > gpurun_out/r02g_bench_n8_iterate10.json 2> gpurun_out/r02g_bench_n8_iterate10.err
tail -4 gpurun_out/r02g_worker8.log
for f in gpurun_out/r02g_bench_n8.json gpurun_out/r02g_bench_n4.json gpurun_out/r02g_bench_n8_nccl.json gpurun_out/r02g_bench_n8_point1000.json gpurun_out/r02g_bench_n8_iterate10.json; do echo $f; head -c 300 $f; echo; done
tail -3 gpurun_out/r02g_bench_n8.err
