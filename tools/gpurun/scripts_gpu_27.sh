#!/bin/bash
# single-CTA two-level plans: dense staging + in-place parking vs strided row reads
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
SZ="8192:1:0:1 1536:1:0:1 3072:1:0:1 6144:1:0:1 2560:1:0:1 5120:1:0:1 10240:1:0:1 7680:1:0:1 9216:1:0:1 12288:1:0:1 16384:0:0:1 8192:1:1:1 3072:1:0:1:d"
echo "== dense staging"; timeout 200 python tools/time_cases.py $SZ
echo "== strided rows"; PFFFT_B200_SPLIT_DENSE=0 timeout 200 python tools/time_cases.py $SZ
echo "== parity"; timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "ordered_parity or inplace or structural or selection" 2>&1 | tail -n 4
