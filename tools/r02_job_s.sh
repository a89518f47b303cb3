#!/bin/bash
# 1-GPU job: end-of-round check: smoke(), the whole GPU suite, the driver's bench command, the reference arm (short)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02s_smoke.log 2>&1
tail -n 2 gpurun_out/r02s_smoke.log
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02s_pytest_gpu.log 2>&1
tail -n 4 gpurun_out/r02s_pytest_gpu.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02s_bench_n1.json 2> gpurun_out/r02s_bench_n1.err
echo "bench rc=$?"
RTB_BENCH_REFERENCE_SECONDS=40 timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r02s_bench_reference.json 2> gpurun_out/r02s_bench_reference.err
echo "reference rc=$?"; head -c 400 gpurun_out/r02s_bench_reference.json; echo
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02s_bench_n1.json").read().strip().splitlines()[-1])
print("ms %.2f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], d["roofline"]["kernel"], d["clocks"], "e2e ms %.2f"%d["e2e"]["ms_per_step"], d["parity"]["ok"], d["cpu_baseline"]["value"])
for k,v in d.get("secondary",{}).items(): print("  ",k,"ms %.2f"%v["ms_per_step"], "frac", v.get("roofline_frac"), v.get("roofline_kernel"), v.get("traffic"), v.get("parity",{}).get("ok"))
PY
