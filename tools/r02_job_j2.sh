#!/bin/bash
# 1-GPU job: ncu --set full of mid-sweep layer launches of the dominant kernel (default bench workload)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_cell2 --launch-skip 900 -c 4 -f -o gpurun_out/r02u_sweep_cell2 python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r02u_ncu_full.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/*.ncu-rep
