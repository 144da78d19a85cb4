#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_full.log
tail -n 6 gpurun_out/pytest_full.log
timeout 900 python bench_configs.py 2> gpurun_out/configs.err | cut -c1-330
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cta_fft -s 2 -c 1 -o gpurun_out/prof_c3_r01b python tools/prof_case.py 4096 0 17 0 > gpurun_out/prof_c3.log 2>&1
