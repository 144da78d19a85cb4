#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python bench_configs.py --sweep 2> gpurun_out/configs.err | cut -c1-400
