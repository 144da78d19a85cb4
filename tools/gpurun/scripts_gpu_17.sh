#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_full.log
tail -n 4 gpurun_out/pytest_full.log
python tools/time_cases.py 16:1:0:1 96:1:0:1 960:1:0:1 4000:1:0:1 64:0:0:1 256:0:0:1 512:0:0:1 160:0:0:1 12000:1:0:1 96:1:0:1:d 480:0:1:1
