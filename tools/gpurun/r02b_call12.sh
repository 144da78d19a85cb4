#!/bin/bash
# r02b_call12.sh -- radix kernels of the larger cores: two resident CTAs with spills (radix_b.cu) against one without (radix_c.cu)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
CASES="2000:1:0:1 2592:1:0:1 4000:1:0:1 6000:1:0:1 4000:0:0:1 5184:0:0:1 8000:0:0:1 12000:0:0:1 4000:1:1:1 8000:0:1:1 4000:1:0:0"
echo "== two CTAs (default)"; timeout -k 5 300 python tools/time_cases.py $CASES | tee gpurun_out/r02b_radix_minb.txt
echo "== one CTA, no spills"; PFFFT_B200_RADIX_MINB1=1 timeout -k 5 300 python tools/time_cases.py $CASES | tee -a gpurun_out/r02b_radix_minb.txt
echo "== tests"; PFFFT_B200_RADIX_MINB1=1 timeout -k 5 600 python -m pytest tests/test_radix_gpu.py -m gpu -q -x 2>&1 | tail -n 3
