#!/bin/bash
# r02b_call29.sh -- backward-real pair form where it measured faster only: spot timings + full suite (last call of the round)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout -k 5 300 python tools/time_cases.py 5184:0:1:1 7680:0:1:1 3200:0:1:1 2304:0:1:1 4096:0:1:1 | tee gpurun_out/r02b_radix_bwd_pairs2.txt
timeout -k 5 2400 python -m pytest tests -m gpu -q 2>&1 | tail -n 5 | tee gpurun_out/r02b_suite_final3.txt
