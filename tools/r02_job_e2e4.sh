#!/bin/bash
# 4-GPU job: phase times of the host-buffer step (RTB200_TIMING=1) on rank 0
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
RTB200_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 4 --steps 5 --warmup 5 --no-cpu-baseline > gpurun_out/r02_e2e4.json 2> gpurun_out/r02_e2e4.err
grep "rtb200 timing" gpurun_out/r02_e2e4.err | tail -n 8
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_e2e4.json").read().strip().splitlines()[-1]); print("N=4 resident %.2f e2e %.2f"%(d["ms_per_step"], d["e2e"]["ms_per_step"]))
PY
