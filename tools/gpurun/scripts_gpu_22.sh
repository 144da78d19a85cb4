#!/bin/bash
# ncu captures of the cluster kernel (strided rows / DSMEM rows), N=16384 complex
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
PFFFT_B200_CLUSTER_SCATTER=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cluster -s 2 -c 1 -f -o gpurun_out/cl4_strided python tools/prof_case.py 16384 1 12 0 > gpurun_out/ncu_cl4_strided.log 2>&1
PFFFT_B200_CLUSTER_SCATTER=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cluster -s 2 -c 1 -f -o gpurun_out/cl4_dsmem python tools/prof_case.py 16384 1 12 0 > gpurun_out/ncu_cl4_dsmem.log 2>&1
tail -n 3 gpurun_out/ncu_cl4_strided.log gpurun_out/ncu_cl4_dsmem.log
ls -la gpurun_out
