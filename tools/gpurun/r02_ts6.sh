#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_ts_gpu.py tests/test_radix_gpu.py -q 2>&1 | tail -n 8
T="timeout 300 python tools/time_cases.py"
for SH in 0 1; do echo "== ts v5 shape $SH"; PFFFT_B200_TS=1 PFFFT_B200_TS_SHAPE=$SH $T 8192:1:0:1 16384:1:0:1 65536:1:0:1 1048576:1:0:1 131072:0:0:1; done
echo "== radix (default plans first, then radix)"
C="16:1:0:1 96:0:0:1 160:0:0:1 288:0:0:1 480:0:0:1 400:1:0:1 800:0:0:1 864:0:0:1 2592:1:0:1 2592:0:0:1 4000:1:0:1 4000:0:0:1 12000:1:0:1 12000:0:0:1"
$T $C
PFFFT_B200_RADIX=1 $T $C
echo "== ncu ts 65536 shape0"
PFFFT_B200_TS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ts -s 2 -c 1 -f -o gpurun_out/r02_ts6_65536 python tools/prof_case.py 65536 1 11 0 > gpurun_out/ncu_ts6.log 2>&1; tail -n 1 gpurun_out/ncu_ts6.log
