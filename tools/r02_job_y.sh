#!/bin/bash
# why is the point workload slower as a `secondary` line than on its own?
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python bench.py --workload "point-128^3-amr-100src" --no-secondary --no-cpu-baseline --steps 3 --warmup 5 > gpurun_out/r02y_pointA.json 2>/dev/null
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --secondary "point-128^3-amr-100src" > gpurun_out/r02y_pointB.json 2>/dev/null
timeout 600 python bench.py --workload "diffuse-64^3-uniform-192dir" --steps 2 --warmup 3 --no-cpu-baseline --secondary "point-128^3-amr-100src" > gpurun_out/r02y_pointC.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02y_pointA.json").read().strip().splitlines()[-1]); print("A standalone: resident %.2f e2e %.2f"%(d["ms_per_step"], d["e2e"]["ms_per_step"]))
for f in ("B","C"):
    d=json.loads(open("gpurun_out/r02y_point%s.json"%f).read().strip().splitlines()[-1]); s=d["secondary"]["point-128^3-amr-100src"]; print(f, "after", d["config"]["workload"], ": resident %.2f e2e %.2f"%(s["ms_per_step"], s["e2e_ms_per_step"]))
PY
