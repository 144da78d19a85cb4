#!/bin/bash
# tiled Stockham pipeline: first hardware run (tests under a short timeout: a protocol bug would spin), then timings
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== ts tests"; timeout 600 python -m pytest tests/test_ts_gpu.py -x -q 2>&1 | tail -n 8
T="timeout 300 python tools/time_cases.py"
echo "== default plans"; $T 8192:1:0:1 12288:1:0:1 16384:1:0:1 32768:1:0:1 65536:1:0:1 36864:1:0:1 61440:1:0:1 32768:0:0:1 131072:0:0:1
echo "== ts"; PFFFT_B200_TS=1 $T 8192:1:0:1 12288:1:0:1 16384:1:0:1 32768:1:0:1 65536:1:0:1 36864:1:0:1 61440:1:0:1 32768:0:0:1 131072:0:0:1 131072:0:1:1 65536:1:0:0 65536:1:1:0
echo "== ts large"; $T 131072:1:0:1 262144:1:0:1 1048576:1:0:1 4194304:1:0:1 16777216:1:0:1 67108864:1:0:1 589824:1:0:1 384000:1:0:1 2097152:0:0:1
echo "== ts double"; PFFFT_B200_TS=1 $T 16384:1:0:1:d 65536:1:0:1:d 1048576:1:0:1:d
echo "== ts lag sweep 65536"; for L in 0 1 4 10 20; do PFFFT_B200_TS=1 PFFFT_B200_TS_LAG=$L $T 65536:1:0:1; done
echo "== ncu ts 65536"
PFFFT_B200_TS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ts -s 2 -c 1 -f -o gpurun_out/r02_ts_65536 python tools/prof_case.py 65536 1 11 0 > gpurun_out/ncu_ts.log 2>&1; tail -n 1 gpurun_out/ncu_ts.log
