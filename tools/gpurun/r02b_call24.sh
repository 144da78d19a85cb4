#!/bin/bash
# r02b_call24.sh -- pair epilogue restricted to the large three-stage cores: spot timings, full suite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout -k 5 300 python tools/time_cases.py 96:0:0:1 288:0:0:1 480:0:0:1 2304:0:0:1 5120:0:0:1 8000:0:0:1 9600:0:0:1 64:0:0:1:d 256:0:0:1:d 512:0:0:1:d 4096:0:0:1 1024:1:0:1 | tee gpurun_out/r02b_radix_pairs2.txt
timeout -k 5 2400 python -m pytest tests -m gpu -q 2>&1 | tail -n 12 | tee gpurun_out/r02b_suite5.txt
