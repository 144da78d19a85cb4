#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -n 15
T="timeout 300 python tools/time_cases.py"
echo "== radix v2 (pair rotation)"; $T 96:0:0:1 160:0:0:1 288:0:0:1 480:0:0:1 800:0:0:1 864:0:0:1 2592:0:0:1 4000:0:0:1 12000:0:0:1 800:0:1:1 4000:0:1:1 16:1:1:1 400:1:0:0 4000:1:1:1
echo "== defaults large"; $T 20480:1:0:1 36864:1:0:1 61440:1:0:1 65536:1:0:1 131072:0:0:1 16384:1:0:1:d 65536:1:0:1:d 1048576:1:0:1
