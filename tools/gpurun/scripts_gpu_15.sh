#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cta_fft -s 2 -c 1 -o gpurun_out/prof_c3_r01c python tools/prof_case.py 4096 0 17 0 > gpurun_out/prof_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fastconv -s 2 -c 1 -o gpurun_out/prof_c4_r01 python tools/prof_fastconv.py > gpurun_out/prof_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cta_fft -s 2 -c 1 -o gpurun_out/prof_c4096_r01 python tools/prof_case.py 4096 1 16 0 > gpurun_out/prof_c4096.log 2>&1
ls -la gpurun_out/*.ncu-rep
