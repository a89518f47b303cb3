#!/bin/bash
# 1-GPU job: directions per task / cells per thread on the shards of an 8- and 4-GPU run
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python tools/shard_times.py 8 0 dirs_per_task=4 dirs_per_task=3 dirs_per_task=2 cells=1 cells=1,dirs_per_task=4 cells=1,dirs_per_task=3 cells=1,dirs_per_task=2 persistent=1,dirs_per_task=4 persistent=1,dirs_per_task=2 > gpurun_out/r02i_shard8.log 2>&1
timeout 600 python tools/shard_times.py 4 0 cells=1 dirs_per_task=4 cells=1,dirs_per_task=4 > gpurun_out/r02i_shard4.log 2>&1
grep -E "max/mean" gpurun_out/r02i_shard8.log gpurun_out/r02i_shard4.log
