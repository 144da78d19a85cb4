#!/bin/bash
# decimation-in-frequency two-pass plans and DIF cluster kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== cluster tests"; timeout 420 python -m pytest tests/test_cluster_gpu.py -x -q 2>&1 | tail -n 5
T="timeout 120 python tools/time_cases.py"
echo "== two-pass DIF (default)"; PFFFT_B200_CLUSTER=0 $T 16384:1:0:1 32768:1:0:1 65536:1:0:1 36864:1:0:1 131072:0:0:1 16384:1:0:1:d
echo "== two-pass DIT"; PFFFT_B200_CLUSTER=0 PFFFT_B200_SPLIT_DIF=0 $T 16384:1:0:1 32768:1:0:1 65536:1:0:1 36864:1:0:1 131072:0:0:1 16384:1:0:1:d
echo "== cluster DIF"; PFFFT_B200_CLUSTER=all PFFFT_B200_CLUSTER_MODE=2 $T 16384:1:0:1 32768:1:0:1
PFFFT_B200_CLUSTER=all PFFFT_B200_CLUSTER_MODE=2 PFFFT_B200_CLUSTER_R16=16 $T 65536:1:0:1
PFFFT_B200_CLUSTER_8192=1 PFFFT_B200_CLUSTER_MODE=2 $T 8192:1:0:1
echo "== parity on split sizes"; timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "ordered_parity or structural or inplace" 2>&1 | tail -n 4
