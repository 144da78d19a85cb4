#!/bin/bash
# r02b_call27.sh -- pipeline default for double cores 15360 ... 131071 too; full suite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
C="18432:1:0:1:d 20480:1:0:1:d 23040:1:0:1:d 36864:1:0:1:d 49152:1:0:1:d 98304:1:0:1:d 36864:0:0:1:d"
echo "== pipeline"; timeout -k 5 300 python tools/time_cases.py $C | tee gpurun_out/r02b_ts_default_d.txt
echo "== previous"; PFFFT_B200_TS=0 timeout -k 5 300 python tools/time_cases.py $C | tee -a gpurun_out/r02b_ts_default_d.txt
echo "== suite"; timeout -k 5 2400 python -m pytest tests -m gpu -q 2>&1 | tail -n 5 | tee gpurun_out/r02b_suite_final2.txt
