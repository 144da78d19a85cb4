#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -n 8
T="timeout 300 python tools/time_cases.py"
echo "== changed paths"; $T 16384:1:0:1 16384:1:1:1 32768:0:0:1 800:0:1:1 4000:0:1:1 480:0:1:1 2592:0:1:1 12000:0:1:1 800:0:1:0
NCU="timeout 300 ncu --set full --clock-control none --import-source on -s 2 -c 1 -f"
$NCU -k regex:k_c1024_ldg -o gpurun_out/r02_c1024 python tools/prof_case.py 1024 1 18 0 > gpurun_out/ncu_a.log 2>&1; tail -n 1 gpurun_out/ncu_a.log
$NCU -k regex:k_cta_fft -o gpurun_out/r02_c3 python tools/prof_case.py 4096 0 16 0 > gpurun_out/ncu_b.log 2>&1; tail -n 1 gpurun_out/ncu_b.log
$NCU -k regex:k_cta_radix -o gpurun_out/r02_radix_400c python tools/prof_case.py 400 1 20 0 > gpurun_out/ncu_c.log 2>&1; tail -n 1 gpurun_out/ncu_c.log
$NCU -k regex:k_cta_radix -o gpurun_out/r02_radix_4000c python tools/prof_case.py 4000 1 16 0 > gpurun_out/ncu_d.log 2>&1; tail -n 1 gpurun_out/ncu_d.log
$NCU -k regex:k_cta_radix -o gpurun_out/r02_radix_800r python tools/prof_case.py 800 0 19 0 > gpurun_out/ncu_e.log 2>&1; tail -n 1 gpurun_out/ncu_e.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_t2dg -s 4 -c 2 -f -o gpurun_out/r02_t2dg_65536 python tools/prof_case.py 65536 1 10 0 > gpurun_out/ncu_f.log 2>&1; tail -n 1 gpurun_out/ncu_f.log
$NCU -k regex:k_t2d_cluster -o gpurun_out/r02_t2dc_16384 python tools/prof_case.py 16384 1 12 0 > gpurun_out/ncu_g.log 2>&1; tail -n 1 gpurun_out/ncu_g.log
$NCU -k regex:k_warp_mixed -o gpurun_out/r02_wmixed_96 python tools/prof_case.py 96 1 21 0 > gpurun_out/ncu_h.log 2>&1; tail -n 1 gpurun_out/ncu_h.log
echo "== sanitizer"
for tool in memcheck racecheck; do timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_cases.py > gpurun_out/r02_sanitize_$tool.log 2>&1; echo "$tool rc=$?"; tail -n 3 gpurun_out/r02_sanitize_$tool.log; done
