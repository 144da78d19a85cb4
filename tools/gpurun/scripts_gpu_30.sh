#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 40 python -m pytest tests/test_tiled2d_gpu.py tests/test_parity_gpu.py -x -q -k "tiled2d_vs or selection or (ordered_parity and (32768 or 65536))" 2>&1 | tail -n 3
