#!/bin/bash
# cluster kernels (first contact), zreorder/zconvolve roofline, ramped host pipeline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== cluster tests"; timeout 420 python -m pytest tests/test_cluster_gpu.py -x -q 2>&1 | tail -n 15
T="timeout 120 python tools/time_cases.py"
echo "== default plans"; $T 16384:1:0:1 32768:1:0:1 65536:1:0:1 16384:1:1:1 32768:0:0:1 131072:0:0:1 8192:1:0:1
echo "== strided rows"; PFFFT_B200_CLUSTER_SCATTER=0 $T 16384:1:0:1 32768:1:0:1
echo "== 16-CTA clusters, dsmem rows"; PFFFT_B200_CLUSTER_R16=16 $T 65536:1:0:1
echo "== 16-CTA clusters, strided rows"; PFFFT_B200_CLUSTER_R16=16 PFFFT_B200_CLUSTER_SCATTER=0 $T 65536:1:0:1
echo "== 8192 on 2-CTA clusters"; PFFFT_B200_CLUSTER_8192=1 $T 8192:1:0:1; PFFFT_B200_CLUSTER_8192=1 PFFFT_B200_CLUSTER_SCATTER=0 $T 8192:1:0:1
echo "== two-pass baseline"; PFFFT_B200_CLUSTER=0 $T 16384:1:0:1 32768:1:0:1 65536:1:0:1
echo "== single-CTA two-level plans (row-major twiddles)"; $T 1536:1:0:1 3072:1:0:1 6144:1:0:1 7680:1:0:1 9216:1:0:1 12288:1:0:1
echo "== spectral kernels"; timeout 300 python bench_configs.py --only-spectral 2>&1 | cut -c1-220
echo "== bench (ramp on)"; timeout 300 python bench.py --steps 5 --no-cpu 2>&1 | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['e2e'])"
echo "== bench (ramp off)"; PFFFT_B200_RAMP=0 timeout 300 python bench.py --steps 5 --no-cpu 2>&1 | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'])"
echo "== bench e2e batch 2^18"; timeout 300 python bench.py --steps 5 --no-cpu --e2e-batch 262144 2>&1 | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['e2e'])"
echo "== full gpu suite"; timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_full.log; tail -n 6 gpurun_out/pytest_full.log
