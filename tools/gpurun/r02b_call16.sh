#!/bin/bash
# r02b_call16.sh -- twelve new three-stage radix cores (1152 ... 3840) against the plans they replace (PFFFT_B200_RADIX=0)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
C="1152:1:0:1 1200:1:0:1 1280:1:0:1 1440:1:0:1 1600:1:0:1 1728:1:0:1 1920:1:0:1 2304:1:0:1 3200:1:0:1 3456:1:0:1 3600:1:0:1 3840:1:0:1 2304:0:0:1 2400:0:0:1 3200:0:0:1 3840:0:0:1 6400:0:0:1 7680:0:0:1 3200:0:1:1 1920:1:0:0"
echo "== radix"; timeout -k 5 300 python tools/time_cases.py $C | tee gpurun_out/r02b_radix_new.txt
echo "== previous plans"; PFFFT_B200_RADIX=0 timeout -k 5 300 python tools/time_cases.py $C | tee -a gpurun_out/r02b_radix_new.txt
echo "== radix tests"; timeout -k 5 900 python -m pytest tests/test_radix_gpu.py tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -n 3
