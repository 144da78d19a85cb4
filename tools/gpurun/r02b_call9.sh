#!/bin/bash
# r02b_call9.sh -- mirrored backward-real CTA passes + final ts defaults: full GPU suite, timings
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== full gpu suite"; timeout -k 5 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 6 | tee gpurun_out/r02b_suite.txt
echo "== time"; timeout -k 5 600 python tools/time_cases.py 4096:0:1:1 4096:0:0:1 2048:0:1:1 1024:0:1:1 8192:0:1:1 4096:0:1:0 131072:1:0:1 1048576:1:0:1 16777216:1:0:1 67108864:1:0:1 589824:1:0:1 384000:1:0:1 131072:0:0:1 2097152:0:0:1 1048576:1:0:1:d | tee gpurun_out/r02b_time9.txt
echo "== ts opt-in at two-pass sizes"; PFFFT_B200_TS=1 timeout -k 5 300 python tools/time_cases.py 8192:1:0:1 16384:1:0:1 32768:1:0:1 65536:1:0:1 36864:1:0:1 20480:1:0:1 | tee -a gpurun_out/r02b_time9.txt
echo "== configs"; timeout -k 5 600 python bench_configs.py --no-cpu --no-spectral 2>&1 | tee gpurun_out/r02b_configs9.json | cut -c1-300
