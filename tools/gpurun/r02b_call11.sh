#!/bin/bash
# r02b_call11.sh -- ring depth / L2 budget of the interleaved pipeline (write-back of the intermediates), new variant tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== variant tests"; timeout -k 5 600 python -m pytest tests/test_ts_gpu.py -m gpu -q -x -k "opt_in" 2>&1 | tail -n 3
export PFFFT_B200_TS=1
for mb in 8 16 24 48 96; do echo "== RING_MB=$mb"; PFFFT_B200_TS_RING_MB=$mb timeout -k 5 300 python tools/time_cases.py 16384:1:0:1 65536:1:0:1 131072:1:0:1 1048576:1:0:1; done 2>&1 | tee gpurun_out/r02b_ring.txt
for lag in 0 2 3; do echo "== LAG=$lag (2^20, 2^22)"; PFFFT_B200_TS_LAG=$lag timeout -k 5 300 python tools/time_cases.py 1048576:1:0:1 4194304:1:0:1 589824:1:0:1; done 2>&1 | tee -a gpurun_out/r02b_ring.txt
