#!/bin/bash
# multi-GPU validation (run with gpurun --gpus 2 or more): C-ABI multi entry points, C example, bench.py under torchrun
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
N=$(nvidia-smi -L | wc -l)
echo "== multi tests on $N GPUs"; NCCL_DEBUG=WARN timeout 600 python -m pytest tests/test_multi_gpu.py -q 2>&1 | tail -n 6
echo "== C example"; gcc -std=c99 -Iinclude/pffft examples/multi_gpu_c2c.c -Lpffft_b200 -lpffft_b200 -Wl,-rpath,$PWD/pffft_b200 -lm -o /tmp/multi_gpu_c2c && timeout 300 /tmp/multi_gpu_c2c 0 16
echo "== bench torchrun N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/bench_n$N.err; cut -c1-600 gpurun_out/r02_bench_n$N.json; tail -n 3 gpurun_out/bench_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n$N.json')); print(d['value'], d['e2e']['value'], d['config'].get('table_broadcast'))"
