#!/bin/bash
# 1-GPU job: compute-sanitizer memcheck over small calls of every path
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 300 python tools/sanitize_small.py > gpurun_out/r02m_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/r02m_plain.log; exit 1; }
tail -5 gpurun_out/r02m_plain.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/r02m_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|Invalid|Error" gpurun_out/r02m_memcheck.log | head -10
