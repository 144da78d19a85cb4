#!/bin/bash
# r02b_ts.sh -- tiled Stockham pipeline with the cp.async input prefetch: correctness, then prefetch on/off x CTAs per SM
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== ubench"; timeout 120 tools/bin/ubench_fp32 | tee gpurun_out/r02b_ubench.txt
echo "== tests (ts, large n)"; PFFFT_B200_TS=1 timeout 1500 python -m pytest tests/test_ts_gpu.py tests/test_large_n_gpu.py -m gpu -q -x 2>&1 | tail -n 5
CASES="16384:1:0:1 32768:1:0:1 65536:1:0:1 36864:1:0:1 131072:1:0:1 1048576:1:0:1 16777216:1:0:1 131072:0:0:1 65536:1:1:1"
for pre in 1 0; do for minb in 3 2 4; do
  echo "== PRE=$pre MINB=$minb"
  PFFFT_B200_TS=1 PFFFT_B200_TS_PRE=$pre PFFFT_B200_TS_MINB=$minb timeout 600 python tools/time_cases.py $CASES
done; done 2>&1 | tee gpurun_out/r02b_ts_matrix.txt
