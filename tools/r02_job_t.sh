#!/bin/bash
# 2-GPU job: device-group tests (two devices in one process, two processes), bench at N = 2 with the phase times of the host-buffer step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r02t_pytest_multi.log 2>&1
tail -n 4 gpurun_out/r02t_pytest_multi.log
RTB200_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r02t_bench_n2.json 2> gpurun_out/r02t_bench_n2.err
echo "bench rc=$?"
grep "rtb200 timing" gpurun_out/r02t_bench_n2.err | tail -n 4
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02t_bench_n2.json").read().strip().splitlines()[-1])
print("N=2 resident %.2f e2e %.2f"%(d["ms_per_step"], d["e2e"]["ms_per_step"]))
PY
