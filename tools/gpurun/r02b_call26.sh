#!/bin/bash
# r02b_call26.sh -- pipeline as the default for untuned float cores 15360 ... 131071 (were two-launch split plans / global path)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
C="15360:1:0:1 18432:1:0:1 19200:1:0:1 23040:1:0:1 25600:1:0:1 30720:1:0:1 46080:1:0:1 57600:1:0:1 73728:1:0:1 81920:1:0:1 98304:1:0:1 36864:0:0:1 61440:0:0:1"
echo "== pipeline (new default)"; timeout -k 5 300 python tools/time_cases.py $C | tee gpurun_out/r02b_ts_default.txt
echo "== previous plans"; PFFFT_B200_TS=0 timeout -k 5 300 python tools/time_cases.py $C | tee -a gpurun_out/r02b_ts_default.txt
echo "== tests"; timeout -k 5 1200 python -m pytest tests/test_parity_gpu.py tests/test_ts_gpu.py tests/test_large_n_gpu.py -m gpu -q 2>&1 | tail -n 4
