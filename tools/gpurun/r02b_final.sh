#!/bin/bash
# r02b_final.sh -- last call of the round: smoke(), full suite, validator sizes, BASELINE configs, bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== smoke"; timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
echo "== suite"; timeout -k 5 2400 python -m pytest tests -m gpu -q 2>&1 | tail -n 4 | tee gpurun_out/r02b_suite_final.txt
V="16 32 64 96 128 160 192 256 288 384 480 512 576 640 800 864 1024 2048 2592 4000 4096 12000 36864"
C=""; for n in $V; do C="$C $n:1:0:1"; done; for n in $V; do [ $n != 16 ] && C="$C $n:0:0:1"; done
echo "== validator sizes (bench_pffft.c:445), complex then real, forward ordered"; timeout -k 5 600 python tools/time_cases.py $C | tee gpurun_out/r02b_validator_sizes.txt
echo "== configs"; timeout -k 5 900 python bench_configs.py 2>&1 | tee gpurun_out/r02b_configs.json | cut -c1-260
echo "== bench"; timeout -k 5 900 python bench.py | tee gpurun_out/r02b_bench_final.json | cut -c1-200
