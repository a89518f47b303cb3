#!/bin/bash
# 1-GPU job: 32-entry exponential in the sweeps: parity tests, bench at 5 and 20 steps
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_diffuse_gpu.py tests/test_diffuse_amr_gpu.py tests/test_multi_gpu.py tests/test_chemistry.py -m gpu -x -q > gpurun_out/r02q_pytest.log 2>&1
tail -4 gpurun_out/r02q_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r02q_bench_n1_5steps.json 2> gpurun_out/r02q_bench_n1_5steps.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02q_bench_n1.json 2> gpurun_out/r02q_bench_n1.err
echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r02q_bench_n1_5steps.json","gpurun_out/r02q_bench_n1.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "ms %.2f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], d["clocks"], d.get("parity",{}).get("ok"), d.get("parity",{}).get("rel_linf_J_fast_vs_faithful_all_192_directions"))
    for k,v in d.get("secondary",{}).items(): print("  ",k,"ms %.2f"%v["ms_per_step"], v.get("parity",{}).get("ok"))
PY
