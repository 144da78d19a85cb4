#!/bin/bash
# r02b_call15.sh -- balanced radices for the cores 48 / 80, core 720 on the radix family; suite; bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
C="48:1:0:1 80:1:0:1 96:0:0:1 160:0:0:1 96:0:1:1 160:0:1:1 48:1:0:0 96:0:0:0"
echo "== balanced (8x6, 10x8)"; timeout -k 5 300 python tools/time_cases.py $C 720:1:0:1 1440:0:0:1 720:1:1:1 1440:0:1:1 720:1:0:0 | tee gpurun_out/r02b_radix_bal.txt
echo "== round-2 shapes (16x3, 16x5)"; PFFFT_B200_RADIX_BAL=0 timeout -k 5 300 python tools/time_cases.py $C | tee -a gpurun_out/r02b_radix_bal.txt
echo "== 720 on the generic kernel"; PFFFT_B200_RADIX=0 timeout -k 5 300 python tools/time_cases.py 720:1:0:1 1440:0:0:1 | tee -a gpurun_out/r02b_radix_bal.txt
echo "== full gpu suite"; timeout -k 5 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 6 | tee gpurun_out/r02b_suite2.txt
echo "== bench"; timeout -k 5 900 python bench.py | tee gpurun_out/r02b_bench_n1b.json | cut -c1-200
