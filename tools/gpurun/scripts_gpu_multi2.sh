#!/bin/bash
# 2 GPUs: sharded stream convolution (NCCL halo message) and a 2-rank bench.py sanity line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/bench_sharded_conv.py > gpurun_out/sharded_conv_n2.log 2>&1; echo "rc=$?"; grep -E "^\{" gpurun_out/sharded_conv_n2.log || tail -n 15 gpurun_out/sharded_conv_n2.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/bench_sharded_conv.py --no-overlap > gpurun_out/sharded_conv_n2_seq.log 2>&1; grep -E "^\{" gpurun_out/sharded_conv_n2_seq.log || tail -n 15 gpurun_out/sharded_conv_n2_seq.log
timeout 100 python tools/bench_sharded_conv.py > gpurun_out/sharded_conv_n1.log 2>&1; grep -E "^\{" gpurun_out/sharded_conv_n1.log
