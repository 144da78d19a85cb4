#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== cluster + fastconv tests"; timeout 420 python -m pytest tests/test_cluster_gpu.py tests/test_fastconv_gpu.py -x -q 2>&1 | tail -n 5
echo "== bench default"; timeout 300 python bench.py --steps 5 --no-cpu 2>&1 | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['e2e']['value'])"
echo "== bench ZC_OUT"; PFFFT_B200_ZC_OUT=1 timeout 300 python bench.py --steps 5 --no-cpu 2>&1 | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e']['max_abs_roundtrip_err'])"
echo "== bench ZC_OUT e2e 2^18"; PFFFT_B200_ZC_OUT=1 timeout 300 python bench.py --steps 5 --no-cpu --e2e-batch 262144 2>&1 | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['e2e']['value'])"
echo "== configs"; timeout 400 python bench_configs.py --no-cpu --no-spectral 2>&1 | grep -E "C4|C3" | cut -c1-300
echo "== sharded conv, 1 rank"; timeout 200 python tools/bench_sharded_conv.py 2>&1 | tail -n 2
echo "== full gpu suite"; timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_full.log; tail -n 5 gpurun_out/pytest_full.log
