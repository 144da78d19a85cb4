#!/usr/bin/env python
"""sanitize_cases.py -- one small launch of every kernel family, for compute-sanitizer (memcheck / racecheck / initcheck).
    compute-sanitizer --tool racecheck python tools/sanitize_cases.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pffft_b200 as pf
rng = np.random.default_rng(0)
def run(N, tr, dt=np.float32, batch=5):
    per = N if tr == 0 else 2 * N
    x = torch.from_numpy((rng.random((batch, per)) * 2 - 1).astype(dt)).cuda()
    with pf.Setup(N, tr, dt) as s:
        f = s.transform_batch(x, 0, True); z = s.transform_batch(x, 0, False)
        b = s.transform_batch(f, 1, True); bz = s.transform_batch(z, 1, False)
        r = s.zreorder_batch(z, 0); s.zreorder_batch(r, 1)
        acc = torch.zeros_like(z); s.zconvolve_batch(z, z, acc, 0.5, True); s.zconvolve_batch(z, z[0].contiguous(), acc, 0.5, False, True)
        torch.cuda.synchronize()
        err = float((b / N - x).abs().max()); errz = float((bz / N - x).abs().max())
        print("%-6d %-7s %-8s %-18s roundtrip %.1e / %.1e" % (N, "real" if tr == 0 else "cplx", np.dtype(dt).name, s.kernel, err, errz), flush=True)
variant = os.environ.get("PFFFT_B200_C1024", "0")
for N, tr in [(1024, 1), (64, 1), (256, 1), (32, 1), (512, 1), (2048, 1), (4096, 1), (1024, 0), (2048, 0), (4096, 0), (8192, 0),
              (16, 1), (96, 1), (160, 0), (480, 1), (8192, 1), (16384, 0), (12000, 1),
              (64, 0), (192, 0), (512, 0), (960, 0), (1920, 0), (800, 1), (128, 1), (48, 1), (4000, 1)]:
    run(N, tr)
for N, tr in [(720, 1), (1440, 0), (2400, 1), (4608, 1), (7680, 1), (9216, 0), (2592, 1), (8000, 0), (432, 1), (160, 0)]:   # radix family, round 2b
    run(N, tr)
for N, tr in [(1024, 1), (4096, 0), (96, 1), (8192, 1), (48, 1), (384, 1), (800, 0), (2592, 0), (2000, 1)]:
    run(N, tr, np.float64, 3)
# overlap-save: fused (Nfft 1024, 8192) and three-launch (Nfft 256, 16384) paths, tail block, complex modes
for taps, bl, n in [(100, 1024, 5000), (4097, 0, 40000), (31, 0, 3000), (200, 16384, 70000)]:
    x = torch.from_numpy((np.arange(n) % 4093).astype(np.float32)).cuda(); y = torch.zeros(n + 64, device="cuda")
    h = np.array([(-1.0, 1.0, 0.5)[j % 3] for j in range(taps)], np.float32)
    for flags in (0, 1, 17):
        fc = pf.FastConv(h, bl, flags)
        got = fc.apply(x, y, n // (2 if flags & 1 else 1), 1)
        torch.cuda.synchronize(); fc.close()
        print("fastconv taps=%d Nfft=%d flags=%d produced=%d" % (taps, fc.block_len, flags, got), flush=True)
# large cores: the default single/two-launch plans and the tiled Stockham pipeline (2, 3 passes, closing small radix,
# pre-/post-rotation stages, ring recycling with a batch larger than the ring)
for N, tr, batch in [(16384, 1, 3), (32768, 1, 3), (65536, 1, 3), (131072, 0, 2), (36864, 1, 2), (131072, 1, 2), (1 << 20, 1, 1),
                     (384000, 1, 1), (1 << 18, 0, 2)]:
    run(N, tr, np.float32, batch)
os.environ["PFFFT_B200_TS"] = "1"
for N, tr, batch in [(16384, 1, 300), (8192, 0, 50), (65536, 1, 8)]:
    run(N, tr, np.float32, batch)
run(16384, 1, np.float64, 20)
os.environ["PFFFT_B200_TSW"] = "1"                      # warp-sized work items (tsw_kernels.cuh), stage-specialised warps
for N, tr, batch in [(16384, 1, 200), (8192, 0, 30), (65536, 1, 6), (131072, 1, 3)]:
    run(N, tr, np.float32, batch)
del os.environ["PFFFT_B200_TSW"]
os.environ["PFFFT_B200_TS_MINB"] = "3"; os.environ["PFFFT_B200_TS_PRE"] = "1"     # cp.async input prefetch variant
run(16384, 1, np.float32, 100)
del os.environ["PFFFT_B200_TS_MINB"]; del os.environ["PFFFT_B200_TS_PRE"]
del os.environ["PFFFT_B200_TS"]
# streaming push / flush and partitioned convolution (C-ABI entry points of round 2)
n, taps = 30000, 301
x = (np.arange(n) % 4093).astype(np.float32)
h = np.array([(-1.0, 1.0, 0.5)[j % 3] for j in range(taps)], np.float32)
for flags in (0, 1):
    fc = pf.FastConv(h, 1024, flags)
    w = 2 if flags & 1 else 1
    xd = torch.from_numpy(x).cuda(); yd = torch.zeros(n + 64, device="cuda")
    got = 0
    for lo in range(0, n // w, 7001):
        c = min(7001, n // w - lo)
        got += fc.push(xd[lo * w:], c, yd[got * w:], fc.pending + c)
    got += fc.flush(yd[got * w:], fc.pending)
    torch.cuda.synchronize(); fc.close()
    print("stream push/flush flags=%d produced=%d" % (flags, got), flush=True)
pc = pf.PartitionedConv(h, 64)
yd = torch.zeros(n, device="cuda")
print("partitioned produced=%d" % pc.apply(torch.from_numpy(x).cuda(), yd, n), flush=True)
torch.cuda.synchronize(); pc.close()
print("done")
