#!/usr/bin/env python
"""one e2e measurement (host pinned buffers, fwd+inv) for the current PFFFT_B200_CHUNK_MB"""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pffft_b200 as pf
N, eb = 1024, 1 << 16
nbytes = eb * 2 * N * 4
bufs = [pf.lib.pffft_aligned_malloc(nbytes) for _ in range(3)]
arr = [np.ctypeslib.as_array(C.cast(b, C.POINTER(C.c_float)), shape=(eb * 2 * N,)) for b in bufs]
arr[0][:] = np.random.default_rng(0).random(eb * 2 * N, dtype=np.float32)
s = pf.Setup(N, 1)
def step():
    pf.pffftb_transform_batch(s.handle, arr[0], arr[1], eb, 0, 1)
    pf.pffftb_transform_batch(s.handle, arr[1], arr[2], eb, 1, 1)
step(); step()
t0 = time.perf_counter()
for _ in range(8): step()
dt = (time.perf_counter() - t0) / 8
print("chunk_mb=%s  e2e %.2f M FFT/s  (%.1f GB/s per direction)" % (os.environ.get("PFFFT_B200_CHUNK_MB", "32"), 2 * eb / dt / 1e6, 2 * nbytes / dt / 1e9))
