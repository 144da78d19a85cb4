#!/bin/bash
# round 1, second session: final measurement set (tests, both bench arms, all configs + sweep, launch list, ncu captures, sanitizer)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
echo "== gpu suite"; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_full.log; tail -n 3 gpurun_out/pytest_full.log
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
echo "== bench reference arm"; timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference_n1.json 2> gpurun_out/bench_reference.err; cut -c1-300 gpurun_out/bench_reference_n1.json
echo "== bench"; timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench.err; cut -c1-1800 gpurun_out/bench_n1.json; tail -n 3 gpurun_out/bench.err
echo "== configs + sweep"; timeout 600 python bench_configs.py > gpurun_out/configs.log 2> gpurun_out/configs.err; grep -c config gpurun_out/configs.log; tail -n 2 gpurun_out/configs.err
echo "== launch list"; timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; grep -c "k_" gpurun_out/launches_bench.csv
echo "== ncu full captures"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_cluster -s 2 -c 1 -f -o gpurun_out/r01b_cluster4 python tools/prof_case.py 16384 1 12 0 > gpurun_out/ncu_a.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_cta_fft -s 2 -c 1 -f -o gpurun_out/r01b_c3 python tools/prof_case.py 4096 0 17 0 > gpurun_out/ncu_b.log 2>&1
tail -n 1 gpurun_out/ncu_a.log gpurun_out/ncu_b.log
echo "== sanitizer (new kernels)"
timeout 150 compute-sanitizer --tool memcheck python tools/sanitize_new.py > gpurun_out/sanitize_memcheck.log 2>&1; tail -n 4 gpurun_out/sanitize_memcheck.log
ls -la gpurun_out | head -40
