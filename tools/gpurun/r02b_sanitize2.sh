#!/bin/bash
# compute-sanitizer over the code added after the full sanitizer run of the round (tools/sanitize_radix_pairs.py)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout -k 5 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_radix_pairs.py > gpurun_out/r02b_sanitize2_$tool.log 2>&1
  echo "$tool rc=$?"; tail -n 3 gpurun_out/r02b_sanitize2_$tool.log
done
