#!/bin/bash
# r02b_call13.sh -- small radix cores: transforms per CTA / resident CTAs (PFFFT_B200_RADIX_VAR) + default check of the larger cores
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
CASES="48:1:0:1 80:1:0:1 144:1:0:1 240:1:0:1 400:1:0:1 432:1:0:1 96:0:0:1 160:0:0:1 288:0:0:1 480:0:0:1 800:0:0:1 864:0:0:1 96:0:1:1 288:0:1:1 800:0:1:1 864:0:1:1"
for v in 0 1 2; do echo "== VAR=$v"; PFFFT_B200_RADIX_VAR=$v timeout -k 5 300 python tools/time_cases.py $CASES; done 2>&1 | tee gpurun_out/r02b_radix_var.txt
echo "== defaults, larger cores"; timeout -k 5 300 python tools/time_cases.py 5184:0:0:1 8000:0:0:1 12000:0:0:1 4000:1:0:1 24000:0:0:1 12000:1:0:1 | tee -a gpurun_out/r02b_radix_var.txt
echo "== radix tests"; timeout -k 5 600 python -m pytest tests/test_radix_gpu.py -m gpu -q -x 2>&1 | tail -n 3
