#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_full.log
tail -n 4 gpurun_out/pytest_full.log
python - <<'PY'
import numpy as np, torch, sys
sys.path.insert(0,'.')
import pffft_b200 as pf
def t(N,tr,lb,dt=torch.float32):
    batch=1<<lb
    per=N if tr==0 else 2*N
    x=torch.rand((batch,per),device='cuda',dtype=dt)*2-1; y=torch.empty_like(x)
    s=pf.Setup(N,tr,np.float32 if dt==torch.float32 else np.float64)
    for _ in range(3): pf.pffftb_transform_batch(s.handle,x,y,batch,0,1)
    torch.cuda.synchronize(); a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): pf.pffftb_transform_batch(s.handle,x,y,batch,0,1)
    b.record(); torch.cuda.synchronize(); ms=a.elapsed_time(b)/10
    es=4 if dt==torch.float32 else 8
    print(N,'real' if tr==0 else 'cplx',str(dt)[6:],s.kernel,'%.3f ms'%ms,'%.0f GB/s'%(2*batch*per*es/ms/1e6), '%.2f of peak'%(2*batch*per*es/ms/1e6/6573.2))
for N in (8192,16384,65536): t(N,1,30-int(np.log2(8*N)))
t(16384,0,30-16); t(131072,0,30-19); t(16384,1,29-17,torch.float64); t(262144,1,10)
t(16,1,23); t(96,1,22); t(960,1,19); t(8192,1,16); t(64,0,23); t(256,0,22); t(512,0,21)
PY
