#!/bin/bash
# 1-GPU job: default bench line, then the ncu launch list of the same command
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02u_bench_n1.json 2> gpurun_out/r02u_bench_n1.err || { echo "bench failed"; tail -5 gpurun_out/r02u_bench_n1.err; exit 1; }
head -c 600 gpurun_out/r02u_bench_n1.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02u_launches.csv python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r02u_ncu_launches.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r02u_launches.csv
