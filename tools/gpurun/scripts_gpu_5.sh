#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cta_fft -s 2 -c 1 -o gpurun_out/prof_c3_r01 python tools/prof_case.py 4096 0 17 0 > gpurun_out/prof_c3.log 2>&1
tail -3 gpurun_out/prof_c3.log
