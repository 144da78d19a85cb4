#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== new tests"; timeout 900 python -m pytest tests/test_ts_gpu.py tests/test_streaming_gpu.py tests/test_multi_gpu.py tests/test_baseline_scale_gpu.py tests/test_fastconv_gpu.py -q 2>&1 | tail -n 12
T="timeout 300 python tools/time_cases.py"
echo "== ts v3 shape 0 (2 CTA x 2 staged)"; PFFFT_B200_TS=1 $T 8192:1:0:1 16384:1:0:1 32768:1:0:1 65536:1:0:1 36864:1:0:1 131072:0:0:1 65536:1:0:0
echo "== ts v3 shape 1 (3 CTA x 1 staged)"; PFFFT_B200_TS=1 PFFFT_B200_TS_SHAPE=1 $T 8192:1:0:1 16384:1:0:1 65536:1:0:1
echo "== ts v3 shape 2 (1 CTA x 4 staged)"; PFFFT_B200_TS=1 PFFFT_B200_TS_SHAPE=2 $T 16384:1:0:1 65536:1:0:1
echo "== ts large"; $T 131072:1:0:1 1048576:1:0:1 16777216:1:0:1 67108864:1:0:1 384000:1:0:1
echo "== ts double shape0/1"; PFFFT_B200_TS=1 $T 16384:1:0:1:d 65536:1:0:1:d 1048576:1:0:1:d; PFFFT_B200_TS=1 PFFFT_B200_TS_SHAPE=1 $T 16384:1:0:1:d 65536:1:0:1:d
echo "== ncu ts 65536"
PFFFT_B200_TS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ts -s 2 -c 1 -f -o gpurun_out/r02_ts3_65536 python tools/prof_case.py 65536 1 11 0 > gpurun_out/ncu_ts3.log 2>&1; tail -n 1 gpurun_out/ncu_ts3.log
