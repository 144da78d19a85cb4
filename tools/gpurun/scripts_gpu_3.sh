#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_full.log
tail -n 5 gpurun_out/pytest_full.log
timeout 900 python bench_configs.py --sweep 2> gpurun_out/configs.err
timeout 300 python bench.py --steps 10 --no-cpu 2>&1 | cut -c1-900
