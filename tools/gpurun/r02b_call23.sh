#!/bin/bash
# r02b_call23.sh -- forward real on the radix kernels with the in-register pair epilogue (radix_last_pairs)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
C="96:0:0:1 160:0:0:1 288:0:0:1 480:0:0:1 800:0:0:1 1440:0:0:1 2304:0:0:1 2400:0:0:1 2560:0:0:1 2880:0:0:1 3200:0:0:1 3456:0:0:1 3840:0:0:1 4000:0:0:1 4608:0:0:1 4800:0:0:1 5120:0:0:1 5760:0:0:1 6400:0:0:1 6912:0:0:1 7200:0:0:1 7680:0:0:1 8000:0:0:1 8640:0:0:1 9600:0:0:1 10240:0:0:1 96:0:0:0 2400:0:0:0 64:0:0:1:d 96:0:0:1:d 256:0:0:1:d 512:0:0:1:d 1920:0:0:1:d 3840:0:0:1:d"
timeout -k 5 400 python tools/time_cases.py $C | tee gpurun_out/r02b_radix_pairs.txt
echo "== tests"; timeout -k 5 900 python -m pytest tests/test_radix_gpu.py -m gpu -q 2>&1 | tail -n 3
