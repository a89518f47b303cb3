#!/bin/bash
# 2-GPU job: device-group tests (two devices in one process, two processes), bench at N = 2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r02t_pytest_multi.log 2>&1
tail -n 4 gpurun_out/r02t_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02t_bench_n2.json 2> gpurun_out/r02t_bench_n2.err
echo "bench rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --steps 10 --warmup 5 --workload combined-64^3-amr3-192dir-64src > gpurun_out/r02t_bench_n2_combined.json 2> gpurun_out/r02t_bench_n2_combined.err
python - <<'PY'
import json
for f in ("gpurun_out/r02t_bench_n2.json","gpurun_out/r02t_bench_n2_combined.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "ms %.2f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], d["roofline"]["kernel"], "e2e ms %.2f"%d["e2e"]["ms_per_step"], d["clocks"])
PY
