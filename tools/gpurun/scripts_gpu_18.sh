#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_full.log
tail -n 4 gpurun_out/pytest_full.log
python tools/time_cases.py 1024:1:0:0 1024:1:1:0 1024:1:0:1 1024:1:1:1
