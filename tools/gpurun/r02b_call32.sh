#!/bin/bash
# r02b_call32.sh -- double radix cores 2160 ... 3840 against the plans they replace; tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
C="2160:1:0:1:d 2304:1:0:1:d 2400:1:0:1:d 2560:1:0:1:d 2592:1:0:1:d 2880:1:0:1:d 3456:1:0:1:d 3600:1:0:1:d 3840:1:0:1:d 5184:0:0:1:d 7680:0:0:1:d"
echo "== radix"; timeout -k 5 300 python tools/time_cases.py $C | tee gpurun_out/r02b_radix_d2.txt
echo "== before"; PFFFT_B200_RADIX_D=0 timeout -k 5 300 python tools/time_cases.py $C | tee -a gpurun_out/r02b_radix_d2.txt
echo "== tests"; timeout -k 5 900 python -m pytest tests/test_radix_gpu.py tests/test_parity_gpu.py -m gpu -q 2>&1 | tail -n 3
