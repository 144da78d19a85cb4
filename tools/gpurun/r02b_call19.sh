#!/bin/bash
# r02b_call19.sh -- 2560 / 5120 / 7680 / 9216: three-stage radix kernels against the one-CTA split kernels / tiled plans
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
C="2560:1:0:1 5120:1:0:1 7680:1:0:1 9216:1:0:1 5120:0:0:1 10240:0:0:1 15360:0:0:1 18432:0:0:1 7680:1:1:1 9216:1:0:0"
echo "== radix"; timeout -k 5 300 python tools/time_cases.py $C | tee gpurun_out/r02b_radix_big.txt
echo "== previous"; PFFFT_B200_RADIX_BIG=0 timeout -k 5 300 python tools/time_cases.py $C | tee -a gpurun_out/r02b_radix_big.txt
echo "== tests"; timeout -k 5 900 python -m pytest tests/test_radix_gpu.py -m gpu -q -x 2>&1 | tail -n 3
