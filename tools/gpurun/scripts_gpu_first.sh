#!/bin/bash
# first GPU contact: smoke, a fast subset of parity tests, then per-variant kernel timing
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | head -20 >> gpurun_out/host.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
for v in 0 1 2 3 4; do
  PFFFT_B200_C1024=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
done
tail -5 gpurun_out/smoke.log gpurun_out/pytest.log; cat gpurun_out/bench_v*.json
