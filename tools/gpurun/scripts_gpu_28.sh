#!/bin/bash
# cluster shapes 4x2 / 4x4, L2-prefetch row mode, row-major combine twiddles
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== cluster tests"; timeout 420 python -m pytest tests/test_cluster_gpu.py -x -q 2>&1 | tail -n 4
T="timeout 120 python tools/time_cases.py"
echo "== two-pass (row-major combine twiddles)"; PFFFT_B200_CLUSTER=0 $T 16384:1:0:1 32768:1:0:1 65536:1:0:1 36864:1:0:1 131072:0:0:1
echo "== 16384"; $T 16384:1:0:1; PFFFT_B200_CLUSTER_MODE=2 $T 16384:1:0:1
echo "== 32768"; PFFFT_B200_CLUSTER=all PFFFT_B200_CLUSTER_SHAPE=4x2 $T 32768:1:0:1; PFFFT_B200_CLUSTER=all PFFFT_B200_CLUSTER_SHAPE=4x2 PFFFT_B200_CLUSTER_MODE=2 $T 32768:1:0:1; PFFFT_B200_CLUSTER=all PFFFT_B200_CLUSTER_MODE=2 $T 32768:1:0:1
echo "== 65536"; PFFFT_B200_CLUSTER=all PFFFT_B200_CLUSTER_SHAPE=4x4 $T 65536:1:0:1; PFFFT_B200_CLUSTER=all $T 65536:1:0:1; PFFFT_B200_CLUSTER=all PFFFT_B200_CLUSTER_MODE=2 $T 65536:1:0:1
