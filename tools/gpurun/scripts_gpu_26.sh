#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_streaming_gpu.py tests/test_cluster_gpu.py -x -q 2>&1 | tail -n 12
