#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_full.log
tail -n 4 gpurun_out/pytest_full.log
timeout 900 python bench_configs.py 2> gpurun_out/configs.err | grep -E "C3|C4" | cut -c1-330
