#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_r01.json 2> gpurun_out/bench_ref_r01.err
timeout 1500 python bench_configs.py --sweep > gpurun_out/configs_lines.txt 2> gpurun_out/configs.err; echo "configs rc=$?"
cut -c1-400 gpurun_out/bench_r01.json; cut -c1-300 gpurun_out/bench_ref_r01.json
python - <<'PY'
import json
for l in open('gpurun_out/configs_lines.txt'):
    d=json.loads(l)
    if 'frac_of_hbm_peak' in d and 'N' in d: print("%-40s N=%-7d %-7s %-16s %7.0f GB/s %.2f"%(d['config'][:40],d['N'],d['transform'],d['kernel'],d['gbs_algorithmic'],d['frac_of_hbm_peak']))
    else: print(l.strip()[:200])
PY
