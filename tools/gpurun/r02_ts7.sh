#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T="timeout 300 python tools/time_cases.py"
for L in 19 28 36 48 64; do echo "== ts v5 lag $L"; PFFFT_B200_TS=1 PFFFT_B200_TS_LAG=$L $T 16384:1:0:1 65536:1:0:1; done
for L in 1 2 3; do echo "== ts v5 2^20 lag $L"; PFFFT_B200_TS_LAG=$L $T 1048576:1:0:1; done
for L in 36; do echo "== ts v5 prefetch lag $L"; PFFFT_B200_TS=1 PFFFT_B200_TS_SHAPE=1 PFFFT_B200_TS_LAG=$L $T 16384:1:0:1 65536:1:0:1; done
echo "== real 131072 lag 36"; PFFFT_B200_TS=1 PFFFT_B200_TS_LAG=36 $T 131072:0:0:1 65536:1:0:0
