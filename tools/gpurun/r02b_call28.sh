#!/bin/bash
# r02b_call28.sh -- backward real with the pre-rotation in registers (radix_first_pairs): A/B against a -DRADIX_NO_PAIRS=1 build; tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
C="2304:0:1:1 2560:0:1:1 3200:0:1:1 3840:0:1:1 4608:0:1:1 5120:0:1:1 5184:0:1:1 6400:0:1:1 7680:0:1:1 9600:0:1:1 10240:0:1:1 2304:0:1:0 512:0:1:1:d 1920:0:1:1:d 3840:0:1:1:d"
echo "== pairs"; timeout -k 5 300 python tools/time_cases.py $C | tee gpurun_out/r02b_radix_bwd_pairs.txt
echo "== no pairs"; PFFFT_B200_LIB=$PWD/pffft_b200/libpffft_b200_nopairs.so timeout -k 5 300 python tools/time_cases.py $C | tee -a gpurun_out/r02b_radix_bwd_pairs.txt
echo "== tests"; timeout -k 5 900 python -m pytest tests/test_radix_gpu.py -m gpu -q 2>&1 | tail -n 3
