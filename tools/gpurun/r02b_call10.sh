#!/bin/bash
# r02b_call10.sh -- interleaved CTA pipeline restored: suite, timings, compute-sanitizer, ncu raw pages (CSV, small), bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== full gpu suite"; timeout -k 5 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 6 | tee gpurun_out/r02b_suite.txt
echo "== tsw tests"; PFFFT_B200_TS=1 PFFFT_B200_TSW=1 timeout -k 5 600 python -m pytest tests/test_ts_gpu.py -m gpu -q -x 2>&1 | tail -n 3 | tee -a gpurun_out/r02b_suite.txt
echo "== time"; timeout -k 5 600 python tools/time_cases.py 131072:1:0:1 1048576:1:0:1 16777216:1:0:1 67108864:1:0:1 589824:1:0:1 384000:1:0:1 2097152:0:0:1 1048576:1:0:1:d 16384:1:0:1:d | tee gpurun_out/r02b_time10.txt
echo "== ts opt-in at two-pass sizes"; PFFFT_B200_TS=1 timeout -k 5 300 python tools/time_cases.py 8192:1:0:1 16384:1:0:1 32768:1:0:1 65536:1:0:1 36864:1:0:1 | tee -a gpurun_out/r02b_time10.txt
echo "== sanitizer"
for tool in memcheck racecheck; do
  timeout -k 5 1200 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_cases.py > gpurun_out/r02b_sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; tail -n 3 gpurun_out/r02b_sanitize_$tool.log
done
cap() { # name regex args...
  name=$1; regex=$2; shift 2
  timeout -k 5 300 ncu --set full --clock-control none -k "regex:$regex" -s 2 -c 1 -f -o /tmp/$name python tools/prof_case.py "$@" > /tmp/$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null; echo "$name: $(tail -n 1 /tmp/$name.log)"
}
PFFFT_B200_TS=1 cap r02b_ncu_ts_65536 k_ts_pipeline 65536 1 11 0
cap r02b_ncu_ts_2p20 k_ts_pipeline 1048576 1 7 0
cap r02b_ncu_c3_bwd k_cta_fft 4096 0 16 1
cap r02b_ncu_zreorder k_zreorder 1024 1 17 0 1 zreorder
cap r02b_ncu_zconvolve k_zconvolve 1024 1 17 0 1 zconvolve
cap r02b_ncu_smem_720 k_smem_fft 720 1 17 0
cap r02b_ncu_wmixed_960 k_warp_mixed 960 1 17 0
cap r02b_ncu_c1024_z k_c1024 1024 1 17 0 0
echo "== bench"; timeout -k 5 900 python bench.py | tee gpurun_out/r02b_bench_n1.json | cut -c1-300
echo "== bench reference"; timeout -k 5 900 python bench.py --impl reference | tee gpurun_out/r02b_bench_reference_n1.json | cut -c1-300
