#!/usr/bin/env python
"""sanitize_radix_pairs.py -- the in-register pair forms of the radix kernels (radix_last_pairs / radix_first_pairs) and the newest
cores, for compute-sanitizer:   compute-sanitizer --tool racecheck python tools/sanitize_radix_pairs.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pffft_b200 as pf
rng = np.random.default_rng(0)
def run(N, tr, dt=np.float32, batch=4):
    per = N if tr == 0 else 2 * N
    x = torch.from_numpy((rng.random((batch, per)) * 2 - 1).astype(dt)).cuda()
    with pf.Setup(N, tr, dt) as s:
        f = s.transform_batch(x, 0, True); z = s.transform_batch(x, 0, False)
        b = s.transform_batch(f, 1, True); bz = s.transform_batch(z, 1, False)
        torch.cuda.synchronize()
        err = float((b / N - x).abs().max()); errz = float((bz / N - x).abs().max())
        print("%-6d %-7s %-8s %-18s roundtrip %.1e / %.1e" % (N, "real" if tr == 0 else "cplx", np.dtype(dt).name, s.kernel, err, errz), flush=True)
# forward pairs (three-stage cores >= 1024), backward pairs (first radix <= 10 or core >= 3840), special rows p = 0 and p = M/2
for N in (2304, 2592, 4000, 5184, 7680, 9600, 10240, 20480, 24000, 28800):
    run(N, 0)
for N in (9600, 13824, 14400, 720, 48):
    run(N, 1)
for N in (512, 1920, 3840):
    run(N, 0, np.float64, 3)
os.environ["PFFFT_B200_TS"] = "1"
run(15360, 1, np.float32, 40); run(36864, 0, np.float32, 20); run(23040, 1, np.float64, 10)
print("done")
