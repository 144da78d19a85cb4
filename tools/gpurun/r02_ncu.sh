#!/bin/bash
# ncu captures of the round: each report is condensed to text on the box (tools/ncu_summary.py + top stall lines) and the
# big .ncu-rep files are dropped (gpurun copies back at most 64 MiB)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== ts tests (cleaned kernel)"; timeout 900 python -m pytest tests/test_ts_gpu.py tests/test_large_n_gpu.py tests/test_radix_gpu.py -q 2>&1 | tail -n 4
NCU="timeout 300 ncu --set full --clock-control none --import-source on -f"
cap() { # name regex skip count  prof_case-args...
  name=$1; regex=$2; skip=$3; count=$4; shift 4
  $NCU -k "regex:$regex" -s $skip -c $count -o /tmp/$name python tools/prof_case.py "$@" > gpurun_out/ncu_$name.log 2>&1
  python tools/ncu_summary.py /tmp/$name.ncu-rep gpurun_out/${name}.txt; tail -n 1 gpurun_out/ncu_$name.log
}
cap r02_ncu_c1024 k_c1024_ldg 2 1 1024 1 18 0
cap r02_ncu_c3 k_cta_fft 2 1 4096 0 16 0
cap r02_ncu_radix_400c k_cta_radix 2 1 400 1 20 0
cap r02_ncu_radix_4000c k_cta_radix 2 1 4000 1 16 0
cap r02_ncu_radix_800r k_cta_radix 2 1 800 0 19 0
cap r02_ncu_t2dg_65536 k_t2dg 4 2 65536 1 10 0
cap r02_ncu_t2dc_16384 k_t2d_cluster 2 1 16384 1 12 0
cap r02_ncu_wmixed_96 k_warp_mixed 2 1 96 1 21 0
cap r02_ncu_ts_2p20 k_ts_pipeline 2 1 1048576 1 7 0
echo "== launch list of bench.py"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_c1024.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; tail -n 2 gpurun_out/r02_launches_bench_c1024.csv | cut -c1-200
echo "== sanitizer"
for tool in memcheck racecheck; do timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_cases.py > gpurun_out/r02_sanitize_$tool.log 2>&1; echo "$tool rc=$?"; tail -n 2 gpurun_out/r02_sanitize_$tool.log; done
echo "== bench"
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/bench.err; cut -c1-400 gpurun_out/r02_bench_n1.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_n1.json 2>> gpurun_out/bench.err; cut -c1-300 gpurun_out/r02_bench_reference_n1.json
echo "== configs"
timeout 900 python bench_configs.py > gpurun_out/r02_configs.json 2> gpurun_out/configs.err; tail -n 5 gpurun_out/r02_configs.json | cut -c1-300
