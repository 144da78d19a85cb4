#!/bin/bash
# r02b_call5.sh -- ts pipeline without the per-item readiness barrier: tests, timings, ncu at 65536; warp kernels on the refined assignment
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== tests"; PFFFT_B200_TS=1 timeout 1500 python -m pytest tests/test_ts_gpu.py tests/test_large_n_gpu.py -m gpu -q -x 2>&1 | tail -n 3
echo "== tests warp"; timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -n 3
echo "== time"; PFFFT_B200_TS=1 timeout 600 python tools/time_cases.py 16384:1:0:1 32768:1:0:1 65536:1:0:1 36864:1:0:1 131072:1:0:1 1048576:1:0:1 16777216:1:0:1 131072:0:0:1 | tee gpurun_out/r02b_ts_v7.txt
timeout 300 python tools/time_cases.py 128:1:0:1 288:1:0:1 320:1:0:1 96:1:0:1 864:1:0:1 256:1:0:0 | tee gpurun_out/r02b_warp_mixed.txt
export PFFFT_B200_TS=1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_ts" -s 2 -c 1 -f -o gpurun_out/r02b_ts7_65536 python tools/prof_case.py 65536 1 11 0 > gpurun_out/r02b_ts7_65536.log 2>&1; tail -n 2 gpurun_out/r02b_ts7_65536.log
