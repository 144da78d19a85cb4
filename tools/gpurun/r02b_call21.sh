#!/bin/bash
# r02b_call21.sh -- final validation: double radix cores 576 ... 1920 (A/B), full suite, compute-sanitizer, bench (both arms)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
C="576:1:0:1:d 640:1:0:1:d 768:1:0:1:d 800:1:0:1:d 864:1:0:1:d 960:1:0:1:d 1152:1:0:1:d 1280:1:0:1:d 1600:1:0:1:d 1920:1:0:1:d 1920:0:0:1:d 3840:0:0:1:d 1600:0:1:1:d"
echo "== radix_d/e"; timeout -k 5 300 python tools/time_cases.py $C | tee gpurun_out/r02b_radix_e.txt
echo "== generic / split"; PFFFT_B200_RADIX_D=0 timeout -k 5 300 python tools/time_cases.py $C | tee -a gpurun_out/r02b_radix_e.txt
echo "== full gpu suite"; timeout -k 5 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 6 | tee gpurun_out/r02b_suite3.txt
echo "== sanitizer"
for tool in memcheck racecheck; do
  timeout -k 5 1500 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_cases.py > gpurun_out/r02b_sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; tail -n 3 gpurun_out/r02b_sanitize_$tool.log
done
echo "== bench"; timeout -k 5 900 python bench.py | tee gpurun_out/r02b_bench_n1c.json | cut -c1-200
echo "== bench reference"; timeout -k 5 900 python bench.py --impl reference | tee gpurun_out/r02b_bench_reference_n1c.json | cut -c1-200
echo "== launches"; timeout -k 5 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches_bench_c1024.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; tail -n 3 gpurun_out/r02b_launches_bench_c1024.csv
