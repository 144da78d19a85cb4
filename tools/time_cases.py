#!/usr/bin/env python
"""time_cases.py N:tr:dir:ordered[:f64] ... -- event-timed GB/s for a list of cases on ~1-2 GiB working sets"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pffft_b200 as pf
for spec in sys.argv[1:]:
    f = spec.split(":")
    N, tr, d, ordered = int(f[0]), int(f[1]), int(f[2]), int(f[3])
    dt = torch.float64 if len(f) > 4 else torch.float32
    es = 8 if dt == torch.float64 else 4
    per = N if tr == 0 else 2 * N
    batch = max(1, (1 << 30) // (per * es))
    x = torch.rand((batch, per), device="cuda", dtype=dt) * 2 - 1
    y = torch.empty_like(x)
    s = pf.Setup(N, tr, np.float32 if es == 4 else np.float64)
    for _ in range(3): pf.pffftb_transform_batch(s.handle, x, y, batch, d, ordered)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): pf.pffftb_transform_batch(s.handle, x, y, batch, d, ordered)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    gbs = 2 * batch * per * es / ms / 1e6
    print("%-22s %-16s %8.3f ms %7.0f GB/s  %.2f of peak" % (spec, s.kernel, ms, gbs, gbs / 6573.2), flush=True)
    s.close(); del x, y
