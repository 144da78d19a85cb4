#!/bin/bash
# r02b_call20.sh -- double-precision radix cores: tests, timings against the generic kernel (PFFFT_B200_RADIX_D=0); full suite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== tests"; timeout -k 5 900 python -m pytest tests/test_radix_gpu.py -m gpu -q -x -k double 2>&1 | tail -n 3
C="16:1:0:1:d 32:1:0:1:d 64:1:0:1:d 96:1:0:1:d 128:1:0:1:d 160:1:0:1:d 192:1:0:1:d 256:1:0:1:d 288:1:0:1:d 384:1:0:1:d 480:1:0:1:d 2000:1:0:1:d 64:0:0:1:d 96:0:0:1:d 256:0:0:1:d 512:0:0:1:d 800:0:0:1:d 2592:0:0:1:d 192:0:1:1:d"
echo "== radix_d"; timeout -k 5 300 python tools/time_cases.py $C | tee gpurun_out/r02b_radix_d.txt
echo "== generic"; PFFFT_B200_RADIX_D=0 timeout -k 5 300 python tools/time_cases.py $C | tee -a gpurun_out/r02b_radix_d.txt
echo "== full gpu suite"; timeout -k 5 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 6 | tee gpurun_out/r02b_suite3.txt
