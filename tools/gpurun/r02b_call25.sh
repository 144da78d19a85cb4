#!/bin/bash
# r02b_call25.sh -- radix cores 9600 ... 14400 against the split plans; tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
C="9600:1:0:1 10800:1:0:1 11520:1:0:1 12960:1:0:1 13824:1:0:1 14400:1:0:1 12000:1:0:1 19200:0:0:1 28800:0:0:1"
echo "== radix"; timeout -k 5 300 python tools/time_cases.py $C | tee gpurun_out/r02b_radix_14400.txt
echo "== previous"; PFFFT_B200_RADIX=0 timeout -k 5 300 python tools/time_cases.py $C | tee -a gpurun_out/r02b_radix_14400.txt
echo "== tests"; timeout -k 5 900 python -m pytest tests/test_radix_gpu.py tests/test_parity_gpu.py -m gpu -q 2>&1 | tail -n 4
