#!/bin/bash
# r02b_call31.sh -- 12800 on the pipeline: parity tests (test_parity_gpu covers the size), timing
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout -k 5 200 python tools/time_cases.py 12800:1:0:1 25600:0:0:1 12800:1:0:1:d
timeout -k 5 900 python -m pytest tests/test_parity_gpu.py tests/test_ts_gpu.py -m gpu -q 2>&1 | tail -n 3
