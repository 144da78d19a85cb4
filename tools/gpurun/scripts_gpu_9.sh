#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
python tools/pcie_probe.py
for mb in 4 8 16 32 64 128; do PFFFT_B200_CHUNK_MB=$mb python tools/e2e_probe.py; done
numactl -H 2>/dev/null | head -5; nvidia-smi topo -m 2>/dev/null | head -12
python bench.py --steps 5 --cpu-seconds 6 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['cpu_baseline']); print(d['clocks']); print(d['e2e']['value'])"
