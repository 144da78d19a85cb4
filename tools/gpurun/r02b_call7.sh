#!/bin/bash
# r02b_call7.sh -- stage-specialised workers (ts and tsw): correctness, timings, ring-depth sweep, ncu
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PFFFT_B200_TS=1
echo "== tests tsw"; timeout -k 5 600 python -m pytest tests/test_ts_gpu.py tests/test_large_n_gpu.py -m gpu -q -x 2>&1 | tail -n 6
echo "== tests ts"; PFFFT_B200_TSW=0 timeout -k 5 600 python -m pytest tests/test_ts_gpu.py tests/test_large_n_gpu.py -m gpu -q -x 2>&1 | tail -n 6
CASES="16384:1:0:1 32768:1:0:1 65536:1:0:1 131072:1:0:1 1048576:1:0:1 16777216:1:0:1"
echo "== tsw default"; timeout -k 5 300 python tools/time_cases.py $CASES 8192:1:0:1 131072:0:0:1 65536:1:0:0 | tee gpurun_out/r02b_stage.txt
echo "== ts default"; PFFFT_B200_TSW=0 timeout -k 5 300 python tools/time_cases.py $CASES 36864:1:0:1 131072:0:0:1 589824:1:0:1 | tee -a gpurun_out/r02b_stage.txt
for lag in 2 6 12; do echo "== LAG=$lag tsw"; PFFFT_B200_TS_LAG=$lag timeout -k 5 300 python tools/time_cases.py 16384:1:0:1 65536:1:0:1 1048576:1:0:1; echo "== LAG=$lag ts"; PFFFT_B200_TSW=0 PFFFT_B200_TS_LAG=$lag timeout -k 5 300 python tools/time_cases.py 16384:1:0:1 65536:1:0:1 1048576:1:0:1; done 2>&1 | tee -a gpurun_out/r02b_stage.txt
for mb in 3 5; do echo "== tsw MINB=$mb"; PFFFT_B200_TSW_MINB=$mb timeout -k 5 300 python tools/time_cases.py 16384:1:0:1 65536:1:0:1 1048576:1:0:1; done 2>&1 | tee -a gpurun_out/r02b_stage.txt
timeout -k 5 400 ncu --set full --clock-control none --import-source on -k "regex:k_tsw" -s 2 -c 1 -f -o gpurun_out/r02b_tsw2_65536 python tools/prof_case.py 65536 1 11 0 > gpurun_out/r02b_tsw2_65536.log 2>&1; tail -n 2 gpurun_out/r02b_tsw2_65536.log
PFFFT_B200_TSW=0 timeout -k 5 400 ncu --set full --clock-control none --import-source on -k "regex:k_ts" -s 2 -c 1 -f -o gpurun_out/r02b_ts8_65536 python tools/prof_case.py 65536 1 11 0 > gpurun_out/r02b_ts8_65536.log 2>&1; tail -n 2 gpurun_out/r02b_ts8_65536.log
