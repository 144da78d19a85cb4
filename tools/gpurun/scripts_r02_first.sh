#!/bin/bash
# FIRST CALL OF ROUND 2: the plans written after round 1's GPU budget was spent (CPU-stepped only) meet the hardware.
#   1. gated tests: cluster-fused tiled plan, general-radix tiled plan, double tiled plan      (each under its own timeout)
#   2. timings of every opt-in plan beside the current default
#   3. one ncu capture of the two tiled passes (65536) for the tuning that follows
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== default suite (tiled2d, cluster)"; timeout 300 python -m pytest tests/test_tiled2d_gpu.py tests/test_cluster_gpu.py -x -q 2>&1 | tail -n 3
echo "== gated: general radix + double"; PFFFT_B200_TEST_T2D_GENERAL=1 timeout 300 python -m pytest tests/test_tiled2d_gpu.py -x -q -k "general or double" 2>&1 | tail -n 5
echo "== gated: cluster-fused (cluster barriers: short timeout)"; PFFFT_B200_TEST_T2D_CLUSTER=1 timeout 120 python -m pytest tests/test_tiled2d_gpu.py -x -q -k cluster_fused 2>&1 | tail -n 5
T="timeout 120 python tools/time_cases.py"
echo "== tiled two-pass (default) vs cluster-fused"; $T 16384:1:0:1 32768:1:0:1 65536:1:0:1
PFFFT_B200_TILED2D=1 $T 16384:1:0:1
PFFFT_B200_TILED2D=2 $T 16384:1:0:1 32768:1:0:1 65536:1:0:1
echo "== general radix: split plans (default) vs tiled"; $T 7680:1:0:1 9216:1:0:1 12288:1:0:1 20480:1:0:1 24576:1:0:1 36864:1:0:1 40960:1:0:1 49152:1:0:1 61440:1:0:1
PFFFT_B200_TILED2D_GENERAL=1 $T 7680:1:0:1 9216:1:0:1 12288:1:0:1 20480:1:0:1 24576:1:0:1 36864:1:0:1 40960:1:0:1 49152:1:0:1 61440:1:0:1 16384:1:0:1 65536:1:0:1
echo "== double: split (default) vs tiled"; $T 16384:1:0:1:d 32768:1:0:1:d 65536:1:0:1:d
PFFFT_B200_TILED2D_GENERAL=1 $T 16384:1:0:1:d 32768:1:0:1:d 65536:1:0:1:d
echo "== ncu: tiled passes at 65536"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_t2d -s 4 -c 2 -f -o gpurun_out/r02_t2d_65536 python tools/prof_case.py 65536 1 10 0 > gpurun_out/ncu_t2d.log 2>&1; tail -n 1 gpurun_out/ncu_t2d.log
echo "== large N (global path, exact division fix)"; timeout 900 python -m pytest tests/test_large_n_gpu.py -x -q 2>&1 | tail -n 8
