#!/bin/bash
# r02b_packed.sh -- first call after moving the float arithmetic to packed f32x2 instructions:
# issue-rate probe, GPU suite, A/B timings packed vs scalar build (PFFFT_B200_LIB), bench line.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== ubench"; timeout 120 gpurun_out/ubench_fp32 | tee gpurun_out/r02b_ubench.txt
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 8 | tee gpurun_out/r02b_tests.txt
CASES="1024:1:0:1 1024:1:1:1 1024:1:0:0 4096:0:0:1 4096:0:1:1 4096:0:0:0 512:1:0:1 2048:1:0:1 4096:1:0:1 8192:0:0:1 256:1:0:1 96:1:0:1 480:1:0:1 800:1:0:1 960:1:0:1 512:0:0:1 16:1:0:1 400:1:0:1 2592:1:0:1 4000:1:0:1 12000:1:0:1 800:0:0:1 2592:0:0:1 8192:1:0:1 6144:1:0:1 16384:1:0:1 32768:1:0:1 65536:1:0:1 36864:1:0:1 131072:1:0:1 1048576:1:0:1 131072:0:0:1 1024:1:0:1:d 4096:0:0:1:d 144:1:0:1 720:1:0:1"
echo "== time packed"; timeout 900 python tools/time_cases.py $CASES | tee gpurun_out/r02b_time_packed.txt
echo "== time scalar"; PFFFT_B200_LIB=$PWD/pffft_b200/libpffft_b200_scalar.so timeout 900 python tools/time_cases.py $CASES | tee gpurun_out/r02b_time_scalar.txt
echo "== ts opt-in 65536/16384"; PFFFT_B200_TS=1 timeout 300 python tools/time_cases.py 65536:1:0:1 16384:1:0:1 32768:1:0:1 | tee gpurun_out/r02b_time_ts.txt
echo "== bench"; timeout 900 python bench.py | tee gpurun_out/r02b_bench.json
