#!/bin/bash
# r02b_call17.sh -- alternative stage shapes of the larger radix cores (radix_x.cu), forward complex / forward real
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
C="2000:1:0:1 2592:1:0:1 4000:1:0:1 6000:1:0:1 12000:1:0:1 4000:0:0:1 5184:0:0:1 8000:0:0:1 12000:0:0:1 24000:0:0:1"
for a in 0 1 2 3; do echo "== ALT=$a"; PFFFT_B200_RADIX_ALT=$a timeout -k 5 300 python tools/time_cases.py $C; done 2>&1 | tee gpurun_out/r02b_radix_alt.txt
