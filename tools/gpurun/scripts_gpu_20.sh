#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_full.log
tail -n 4 gpurun_out/pytest_full.log
python tools/time_cases.py 96:1:0:1 160:1:0:1 192:1:1:1 288:1:0:1 320:1:0:1 384:1:0:1 480:1:0:1 480:1:0:0 96:1:1:0
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python - <<'PY' 2>&1 | grep -E "SUMMARY|hazard|ok" | head
import sys; sys.path.insert(0,'.')
import numpy as np, torch, pffft_b200 as pf
for N in (96,160,192,288,320,384,480):
    x=torch.rand((7,2*N),device='cuda')*2-1
    with pf.Setup(N,1) as s:
        f=s.transform_batch(x,0,True); z=s.transform_batch(x,0,False); b=s.transform_batch(z,1,False); torch.cuda.synchronize()
        print(N,s.kernel,'ok',float((b/N-x).abs().max()))
PY
