#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_full.log
tail -n 60 gpurun_out/pytest_full.log
