#!/bin/bash
# 1-GPU job: config-3 size point workload (256^3 + 1 level, 1000 sources) with the planned deposition
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 400 python bench.py --workload "point-256^3-amr-1000src" --steps 2 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r02y_bench_n1_point256_1000src.json 2> gpurun_out/r02y_bench_n1_point256_1000src.err
echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02y_bench_n1_point256_1000src.json").read().strip().splitlines()[-1]); r=d["roofline"]
print(d["config"]["workload"], "ms %.1f"%d["ms_per_step"], "value %.3g"%d["value"], "frac %.3f"%r["frac"], "e2e %.1f"%d["e2e"]["ms_per_step"])
PY
