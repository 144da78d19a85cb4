#!/bin/bash
# r02b_call14.sh -- core 432 as two stages (24 x 18) against three (12 x 12 x 3); radix defaults re-checked
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
C="432:1:0:1 864:0:0:1 864:0:1:1 432:1:1:1 432:1:0:0 864:0:0:0"
echo "== 24x18"; timeout -k 5 300 python tools/time_cases.py $C | tee gpurun_out/r02b_radix_432.txt
echo "== 12x12x3"; PFFFT_B200_RADIX_432=3 timeout -k 5 300 python tools/time_cases.py $C | tee -a gpurun_out/r02b_radix_432.txt
echo "== defaults"; timeout -k 5 300 python tools/time_cases.py 48:1:0:1 96:0:0:1 96:0:1:1 160:0:0:1 288:0:0:1 288:0:1:1 800:0:1:1 800:0:0:1 | tee -a gpurun_out/r02b_radix_432.txt
echo "== radix tests"; timeout -k 5 600 python -m pytest tests/test_radix_gpu.py -m gpu -q -x 2>&1 | tail -n 3
