#!/bin/bash
# 1-GPU job: A/B of the sweeps' exponential: 32 entries + r^5 against 64 entries + r^4 (same box, alternating)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
cp radiativetransfer_b200/librtb200.so /tmp/lib32.so
cp radiativetransfer_b200/librtb200_exp64.so /tmp/lib64.so
for rep in 1 2; do
  for v in 32 64; do
    cp /tmp/lib$v.so radiativetransfer_b200/librtb200.so
    timeout 600 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r02r_bench_exp${v}_$rep.json 2> gpurun_out/r02r_bench_exp${v}_$rep.err
  done
done
cp /tmp/lib64.so radiativetransfer_b200/librtb200.so
timeout 900 python -m pytest tests/test_diffuse_gpu.py tests/test_diffuse_amr_gpu.py -m gpu -x -q > gpurun_out/r02r_pytest_exp64.log 2>&1
tail -3 gpurun_out/r02r_pytest_exp64.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02r_bench_exp*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "ms %.2f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], d["clocks"]["sm_mhz"], d.get("parity",{}).get("rel_linf_J_fast_vs_faithful_all_192_directions"))
PY
