#!/usr/bin/env python
"""pcie_probe.py -- ceiling for the host-buffer (e2e) path: pinned H2D / D2H bandwidth alone and concurrently,
then the library's host-pointer batch call at several chunk sizes (PFFFT_B200_CHUNK_MB)."""
import os, sys, time, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
n = 1 << 28                                     # 1 GiB of float32 /4 -> 256 Mi floats = 1 GiB
h_in = torch.empty(n, dtype=torch.float32, pin_memory=True); h_out = torch.empty(n, dtype=torch.float32, pin_memory=True)
d_a = torch.empty(n, dtype=torch.float32, device="cuda"); d_b = torch.empty(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timeit(f, reps=5):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
gb = n * 4 / 1e9
t = timeit(lambda: d_a.copy_(h_in, non_blocking=True)); print("H2D alone   %.1f GB/s" % (gb / t))
t = timeit(lambda: h_out.copy_(d_b, non_blocking=True)); print("D2H alone   %.1f GB/s" % (gb / t))
def both():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
t = timeit(both); print("H2D+D2H concurrently: %.1f GB/s each direction" % (gb / t))
