#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Race reported|Invalid|hazard" gpurun_out/sanitizer_$tool.log | head -12
done
for v in 2 3; do PFFFT_B200_C1024=$v PFFFT_B200_CTA_STAGE=1 timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python tools/sanitize_cases.py > gpurun_out/sanitizer_race_v$v.log 2>&1; echo "== racecheck variant $v rc=$?"; grep -E "RACECHECK SUMMARY|Race reported|hazard" gpurun_out/sanitizer_race_v$v.log | head -6; done
tail -3 gpurun_out/sanitizer_memcheck.log
