#!/bin/bash
# r02b_ts2.sh -- ts pipeline, 4 CTAs per SM: tests, then one ncu --set full of the 65536-point pipeline (batch 2048)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== tests"; PFFFT_B200_TS=1 timeout 1500 python -m pytest tests/test_ts_gpu.py tests/test_large_n_gpu.py -m gpu -q -x 2>&1 | tail -n 5
echo "== tests PRE"; PFFFT_B200_TS=1 PFFFT_B200_TS_MINB=3 PFFFT_B200_TS_PRE=1 timeout 1500 python -m pytest tests/test_ts_gpu.py -m gpu -q -x 2>&1 | tail -n 5
echo "== time PRE"; PFFFT_B200_TS=1 PFFFT_B200_TS_MINB=3 PFFFT_B200_TS_PRE=1 timeout 600 python tools/time_cases.py 16384:1:0:1 32768:1:0:1 65536:1:0:1 1048576:1:0:1
export PFFFT_B200_TS=1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_ts" -s 2 -c 1 -f -o gpurun_out/r02b_ts4_65536 python tools/prof_case.py 65536 1 11 0 > gpurun_out/r02b_ts4_65536.log 2>&1; tail -n 2 gpurun_out/r02b_ts4_65536.log
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_ts" -s 2 -c 1 -f -o gpurun_out/r02b_ts4_2p20 python tools/prof_case.py 1048576 1 7 0 > gpurun_out/r02b_ts4_2p20.log 2>&1; tail -n 2 gpurun_out/r02b_ts4_2p20.log
