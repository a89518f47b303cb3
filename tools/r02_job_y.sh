#!/bin/bash
# 1-GPU job: bench lines of the nested-grid workloads with the final binary
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for w in "diffuse-128^3-amr2-192dir" "iterate10-64^3-amr3-192dir"; do
  f=gpurun_out/r02y_bench_n1_$(echo $w | tr '^' '_').json
  timeout 600 python bench.py --workload "$w" --steps 5 --warmup 3 --no-secondary --cpu-seconds 6 > $f 2>/dev/null
  python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
print(d["config"]["workload"], "ms %.2f"%d["ms_per_step"], "value %.3g"%d["value"], "frac %.3f"%r["frac"], r["kernel"], "e2e %.2f"%d["e2e"]["ms_per_step"], d.get("parity",{}).get("ok"))
PY
done
