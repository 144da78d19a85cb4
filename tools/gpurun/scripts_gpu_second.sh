#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
# launch list of the bench command (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu --batch 262144 --e2e-batch 4096 > gpurun_out/bench_under_ncu.log 2>&1
# one full capture of the top kernel (forward c1024)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_c1024 -s 6 -c 2 -o gpurun_out/prof_c1024_r01 \
   python bench.py --steps 3 --warmup 3 --no-cpu --batch 262144 --e2e-batch 4096 > gpurun_out/bench_under_ncu2.log 2>&1
# real bench with CPU baseline and reference arm
timeout 600 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_r01.json 2> gpurun_out/bench_ref_r01.err
tail -n 30 gpurun_out/pytest.log; cat gpurun_out/bench_r01.json gpurun_out/bench_ref_r01.json
