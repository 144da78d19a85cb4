#!/bin/bash
# 2 GPUs, larger shards: 2^27 samples per GPU (the 2^24-sample step is 0.11 ms and host-launch bound)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/bench_sharded_conv.py --samples-per-gpu 134217728 > gpurun_out/sharded_conv_big_n2.log 2>&1; grep -E "^\{" gpurun_out/sharded_conv_big_n2.log || tail -n 15 gpurun_out/sharded_conv_big_n2.log
timeout 100 python tools/bench_sharded_conv.py --samples-per-gpu 134217728 > gpurun_out/sharded_conv_big_n1.log 2>&1; grep -E "^\{" gpurun_out/sharded_conv_big_n1.log || tail -n 5 gpurun_out/sharded_conv_big_n1.log
