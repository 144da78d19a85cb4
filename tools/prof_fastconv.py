#!/usr/bin/env python
"""a few C4 pffastconv_apply calls on device-resident data, for ncu"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pffft_b200 as pf
n, taps = 1 << 24, 4097
x = torch.from_numpy((np.arange(n) % 4093).astype(np.float32)).cuda(); y = torch.empty(n, device="cuda")
h = np.array([(-1.0, 1.0, 0.5)[j % 3] for j in range(taps)], np.float32)
fc = pf.FastConv(h, 0, 0)
for _ in range(4): fc.apply(x, y, n, 1)
torch.cuda.synchronize()
