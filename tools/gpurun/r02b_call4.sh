#!/bin/bash
# r02b_call4.sh -- (1) ts pipeline after the 64-bit-division fix: lag sweep; (2) warp kernels packed vs scalar per R2;
# (3) C3 / C4 configs on the mixed build vs the all-scalar build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PFFFT_B200_TS=1
CASES="16384:1:0:1 65536:1:0:1 36864:1:0:1 1048576:1:0:1 16777216:1:0:1"
for lag in default 12 24 32 48 64; do
  echo "== ts MINB=4 LAG=$lag"
  if [ $lag == default ]; then timeout 600 python tools/time_cases.py $CASES 32768:1:0:1 131072:1:0:1 131072:0:0:1; else PFFFT_B200_TS_LAG=$lag timeout 600 python tools/time_cases.py $CASES; fi
done 2>&1 | tee gpurun_out/r02b_ts_lag.txt
echo "== ts MINB=3 after the division fix"; PFFFT_B200_TS_MINB=3 timeout 600 python tools/time_cases.py 16384:1:0:1 65536:1:0:1 1048576:1:0:1 | tee -a gpurun_out/r02b_ts_lag.txt
unset PFFFT_B200_TS
W="32:1:0:1 64:1:0:1 128:1:0:1 256:1:0:1 96:1:0:1 160:1:0:1 192:1:0:1 288:1:0:1 320:1:0:1 384:1:0:1 480:1:0:1 576:1:0:1 640:1:0:1 768:1:0:1 864:1:0:1 64:0:0:1 128:0:0:1 256:0:0:1 192:0:0:1 384:0:0:1 640:0:0:1 960:0:0:1 1920:0:0:1 512:0:1:1 960:0:1:1 96:1:1:1 480:1:1:1 256:1:0:0 480:1:0:0 1024:1:1:0"
echo "== warp kernels, all packed"; PFFFT_B200_LIB=$PWD/pffft_b200/libpffft_b200_fastpk.so timeout 900 python tools/time_cases.py $W | tee gpurun_out/r02b_warp_pk.txt
echo "== warp kernels, all scalar"; PFFFT_B200_LIB=$PWD/pffft_b200/libpffft_b200_scalar.so timeout 900 python tools/time_cases.py $W | tee gpurun_out/r02b_warp_sc.txt
echo "== configs mixed build"; timeout 600 python bench_configs.py --no-cpu --no-spectral 2>&1 | tee gpurun_out/r02b_configs_mixed.json | cut -c1-400
echo "== configs scalar build"; PFFFT_B200_LIB=$PWD/pffft_b200/libpffft_b200_scalar.so timeout 600 python bench_configs.py --no-cpu --no-spectral 2>&1 | tee gpurun_out/r02b_configs_scalar.json | cut -c1-400
