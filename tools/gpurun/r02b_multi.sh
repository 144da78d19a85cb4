#!/bin/bash
# r02b_multi.sh -- two GPUs: library-level multi-GPU entry points, the torchrun bench at N = 2 (both arms)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== multi-gpu tests"; timeout -k 5 600 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -n 4
echo "== bench N=2"; timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 | tail -n 1 | tee gpurun_out/r02b_bench_n2.json | cut -c1-400
