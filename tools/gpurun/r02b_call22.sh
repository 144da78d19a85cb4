#!/bin/bash
# r02b_call22.sh -- full GPU suite (no -x) on the final build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout -k 5 2400 python -m pytest tests -m gpu -q 2>&1 | tail -n 25 | tee gpurun_out/r02b_suite4.txt
