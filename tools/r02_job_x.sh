#!/bin/bash
# 8-GPU job, final binary: bench lines at N = 8 and 4 (one process per GPU)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for n in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02x_bench_n$n.json 2> gpurun_out/r02x_bench_n$n.err
done
python - <<'PY'
import json
for f in ("gpurun_out/r02x_bench_n8.json","gpurun_out/r02x_bench_n4.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1]); r=d["roofline"]
    print(f, "ms %.3f"%d["ms_per_step"], "e2e ms %.2f"%d["e2e"]["ms_per_step"], "frac %.3f"%r["frac"], r["kernel"], "kernel min/max %.2f/%.2f"%(r["kernel_ms_min_over_ranks"], r["kernel_ms_max_over_ranks"]))
PY
