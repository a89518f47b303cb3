#!/bin/bash
# 1-GPU job: point workload plain, then ncu --set full of one pass (six pixel-level launches) of the march kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 5 --workload "point-128^3-amr-100src" --no-secondary > gpurun_out/r02n_bench_point128.json 2> gpurun_out/r02n_bench_point128.err || { echo "plain bench failed"; tail -5 gpurun_out/r02n_bench_point128.err; exit 1; }
head -c 300 gpurun_out/r02n_bench_point128.json; echo
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:point_march --launch-skip 42 -c 6 -f -o gpurun_out/r02n_point_march python bench.py --steps 2 --warmup 5 --workload "point-128^3-amr-100src" --no-secondary --no-cpu-baseline > gpurun_out/r02n_ncu.log 2>&1
echo "ncu rc=$?"
