#!/bin/bash
# 8-GPU job: device group at 8 ranks (one process per GPU): worker test, bench lines at N = 8, 4, 2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tests/multi_rank_worker.py > gpurun_out/r02p_worker8.log 2>&1
echo "worker rc=$?" >> gpurun_out/r02p_worker8.log
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02p_bench_n$n.json 2> gpurun_out/r02p_bench_n$n.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 8 --steps 5 --warmup 3 --workload iterate10-256^3-uniform-192dir > gpurun_out/r02p_bench_n8_iterate10.json 2> gpurun_out/r02p_bench_n8_iterate10.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29554 bench.py --gpus 8 --steps 10 --warmup 5 --workload diffuse-128^3-amr2-192dir > gpurun_out/r02p_bench_n8_amr128.json 2> gpurun_out/r02p_bench_n8_amr128.err
tail -n 3 gpurun_out/r02p_worker8.log
for f in gpurun_out/r02p_bench_n8.json gpurun_out/r02p_bench_n4.json gpurun_out/r02p_bench_n2.json gpurun_out/r02p_bench_n8_iterate10.json gpurun_out/r02p_bench_n8_amr128.json; do echo $f; head -c 330 $f | tail -c 130; echo; done
tail -n 3 gpurun_out/r02p_bench_n8_amr128.err
