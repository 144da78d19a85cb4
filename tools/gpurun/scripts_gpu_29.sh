#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_parity_gpu.py -x -q -k "16384 or 32768 or 65536 or selection" 2>&1 | tail -n 2
PFFFT_B200_TEST_TILED2D=1 timeout 60 python -m pytest tests/test_tiled2d_gpu.py -x -q 2>&1 | tail -n 6
PFFFT_B200_TILED2D=1 timeout 60 python tools/time_cases.py 16384:1:0:1 32768:1:0:1 65536:1:0:1 65536:1:1:1
