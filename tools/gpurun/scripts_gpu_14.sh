#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_full.log
tail -n 4 gpurun_out/pytest_full.log
python tools/time_cases.py 4096:1:0:0 4096:1:1:0 4096:0:0:0 4096:0:1:0 4096:0:1:1 8192:0:0:0 8192:0:1:0 8192:0:1:1 2048:1:0:0 1024:1:0:0:d 4096:0:0:0:d
timeout 300 python bench_configs.py --no-cpu 2>/dev/null | grep -E "C3|C4" | cut -c1-260
