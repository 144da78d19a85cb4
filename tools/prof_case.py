#!/usr/bin/env python
"""prof_case.py N transform(0|1) log2batch direction [ordered] [op] -- a few launches of one configuration, for ncu
(op: transform (default) | zreorder | zconvolve)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pffft_b200 as pf
N, tr, lb, d = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
ordered = int(sys.argv[5]) if len(sys.argv) > 5 else 1
per = N if tr == 0 else 2 * N
batch = 1 << lb
x = torch.rand((batch, per), device="cuda") * 2 - 1
y = torch.empty_like(x)
s = pf.Setup(N, tr)
op = sys.argv[6] if len(sys.argv) > 6 else "transform"
st = pf.Setup(N, tr)
for _ in range(4):
    if op == "zreorder": st.zreorder_batch(x, d)
    elif op == "zconvolve": st.zconvolve_batch(x, x[0].contiguous(), y, 0.5, True, True)
    else: pf.pffftb_transform_batch(s.handle, x, y, batch, d, ordered)
torch.cuda.synchronize()
print(s.kernel)
