#!/bin/bash
# r02b_call18.sh -- promoted shapes (2000, 2592, 4000, 6000) and the cores 2160 ... 8000: tests, timings against the plans they replace
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== tests"; timeout -k 5 900 python -m pytest tests/test_radix_gpu.py tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -n 3
C="2000:1:0:1 2592:1:0:1 4000:1:0:1 6000:1:0:1 4000:0:0:1 5184:0:0:1 8000:0:0:1 12000:0:0:1 4000:1:1:1 8000:0:1:1 2592:1:0:0 2160:1:0:1 2400:1:0:1 2880:1:0:1 4320:1:0:1 4608:1:0:1 4800:1:0:1 5184:1:0:1 5760:1:0:1 6400:1:0:1 6912:1:0:1 7200:1:0:1 8000:1:0:1 9600:0:0:1 16000:0:0:1"
echo "== radix"; timeout -k 5 400 python tools/time_cases.py $C | tee gpurun_out/r02b_radix_new2.txt
echo "== previous plans"; PFFFT_B200_RADIX=0 timeout -k 5 400 python tools/time_cases.py 2160:1:0:1 2400:1:0:1 2880:1:0:1 4320:1:0:1 4608:1:0:1 4800:1:0:1 5184:1:0:1 5760:1:0:1 6400:1:0:1 6912:1:0:1 7200:1:0:1 8000:1:0:1 9600:0:0:1 16000:0:0:1 | tee -a gpurun_out/r02b_radix_new2.txt
