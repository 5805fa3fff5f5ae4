#!/bin/bash
# compute-sanitizer over a reduced set of hot-path invocations (tools/sanitize_cases.py).  ONE tool per gpurun call (see
# /opt/skills/guides/B200_PROFILING.md):   gpurun -- 'bash tools/sanitize.sh memcheck'   (memcheck | racecheck | initcheck | synccheck)
tool=${1:-memcheck}
mkdir -p gpurun_out
timeout 300 python tools/sanitize_cases.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/sanitize_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python tools/sanitize_cases.py > gpurun_out/sanitize_$tool.log 2>&1
echo "compute-sanitizer --tool $tool exit code $?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Uninitialized|ok$| ok " gpurun_out/sanitize_$tool.log | tail -25
