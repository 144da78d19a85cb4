#!/bin/bash
# cluster kernels after relaxed arrives + cp.async row prefetch
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== cluster tests"; timeout 420 python -m pytest tests/test_cluster_gpu.py -x -q 2>&1 | tail -n 5
T="timeout 120 python tools/time_cases.py"
echo "== strided rows (prefetched)"; PFFFT_B200_CLUSTER_SCATTER=0 $T 16384:1:0:1 32768:1:0:1 65536:1:0:1 16384:1:1:1
echo "== dsmem rows"; $T 16384:1:0:1 32768:1:0:1
echo "== 16-CTA clusters"; PFFFT_B200_CLUSTER_R16=16 PFFFT_B200_CLUSTER_SCATTER=0 $T 65536:1:0:1; PFFFT_B200_CLUSTER_R16=16 $T 65536:1:0:1
echo "== 8192 on 2-CTA clusters"; PFFFT_B200_CLUSTER_8192=1 PFFFT_B200_CLUSTER_SCATTER=0 $T 8192:1:0:1
PFFFT_B200_CLUSTER_SCATTER=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cluster -s 2 -c 1 -f -o gpurun_out/cl4_strided_b python tools/prof_case.py 16384 1 12 0 > gpurun_out/ncu_cl4_strided_b.log 2>&1
tail -n 2 gpurun_out/ncu_cl4_strided_b.log
