#!/bin/bash
# r02b_call30.sh -- the remaining one-CTA split cores (8192, 10240, 12288) against the pipeline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
C="8192:1:0:1 10240:1:0:1 12288:1:0:1 12800:1:0:1 16384:0:0:1 20480:0:0:1 24576:0:0:1"
echo "== default"; timeout -k 5 300 python tools/time_cases.py $C | tee gpurun_out/r02b_split_vs_ts.txt
echo "== pipeline"; PFFFT_B200_TS=1 timeout -k 5 300 python tools/time_cases.py $C | tee -a gpurun_out/r02b_split_vs_ts.txt
