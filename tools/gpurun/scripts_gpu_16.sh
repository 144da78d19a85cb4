#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_smem_fft -s 2 -c 1 -o gpurun_out/prof_gen960 python tools/prof_case.py 960 1 17 0 > gpurun_out/prof_gen.log 2>&1
