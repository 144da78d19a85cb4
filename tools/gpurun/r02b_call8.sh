#!/bin/bash
# r02b_call8.sh -- stage-specialised workers: ring-depth sweep (the producers of a stage hold W_stage / tiles transforms in flight)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PFFFT_B200_TS=1 PFFFT_B200_TS_RING_MB=100
echo "== tests tsw"; timeout -k 5 600 python -m pytest tests/test_ts_gpu.py -m gpu -q -x 2>&1 | tail -n 4
for lag in 40 75 110 160; do
  echo "== LAG=$lag ts";  PFFFT_B200_TSW=0 PFFFT_B200_TS_LAG=$lag timeout -k 5 300 python tools/time_cases.py 16384:1:0:1 32768:1:0:1 65536:1:0:1
  echo "== LAG=$lag tsw"; PFFFT_B200_TS_LAG=$lag timeout -k 5 300 python tools/time_cases.py 16384:1:0:1 32768:1:0:1 65536:1:0:1
done 2>&1 | tee gpurun_out/r02b_stage2.txt
PFFFT_B200_TSW=0 PFFFT_B200_TS_LAG=40 timeout -k 5 400 ncu --set full --clock-control none --import-source on -k "regex:k_ts" -s 2 -c 1 -f -o gpurun_out/r02b_ts8_65536 python tools/prof_case.py 65536 1 11 0 > gpurun_out/r02b_ts8_65536.log 2>&1; tail -n 2 gpurun_out/r02b_ts8_65536.log
PFFFT_B200_TS_LAG=40 timeout -k 5 400 ncu --set full --clock-control none --import-source on -k "regex:k_tsw" -s 2 -c 1 -f -o gpurun_out/r02b_tsw2_65536 python tools/prof_case.py 65536 1 11 0 > gpurun_out/r02b_tsw2_65536.log 2>&1; tail -n 2 gpurun_out/r02b_tsw2_65536.log
