#!/usr/bin/env python
"""sanitize_new.py -- compute-sanitizer cases for the kernels added in the second session of round 1: cluster plans (every
row mode / shape), row-major-twiddle two-level plans, group-wise zreorder, pipelined host path of pffastconv."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pffft_b200 as pf
rng = np.random.default_rng(0)
KEYS = ("PFFFT_B200_TILED2D", "PFFFT_B200_CLUSTER", "PFFFT_B200_CLUSTER_MODE", "PFFFT_B200_CLUSTER_SHAPE", "PFFFT_B200_CLUSTER_R16", "PFFFT_B200_CLUSTER_8192")
def run(N, tr, env, batch=3):
    for k in KEYS: os.environ.pop(k, None)
    os.environ.update(env)
    per = N if tr == 0 else 2 * N
    x = torch.from_numpy((rng.random((batch, per)) * 2 - 1).astype(np.float32)).cuda()
    with pf.Setup(N, tr) as s:
        f = s.transform_batch(x, 0, True); z = s.transform_batch(x, 0, False)
        b = s.transform_batch(f, 1, True)
        r = s.zreorder_batch(z, 0); z2 = s.zreorder_batch(r, 1)
        torch.cuda.synchronize()
        print("%-6d %-5s %-30s roundtrip %.1e reorder %s" % (N, "real" if tr == 0 else "cplx", s.kernel, float((b / N - x).abs().max()),
                                                              bool(torch.equal(z, z2))), flush=True)
A = {"PFFFT_B200_CLUSTER": "all", "PFFFT_B200_TILED2D": "0"}
run(16384, 1, {}); run(16384, 1, {"PFFFT_B200_CLUSTER_MODE": "1"}); run(32768, 0, {})
run(32768, 1, A); run(32768, 1, dict(A, PFFFT_B200_CLUSTER_MODE="1")); run(32768, 1, dict(A, PFFFT_B200_CLUSTER_SHAPE="4x2"))
run(65536, 1, A); run(65536, 1, dict(A, PFFFT_B200_CLUSTER_R16="16")); run(65536, 1, dict(A, PFFFT_B200_CLUSTER_R16="16", PFFFT_B200_CLUSTER_MODE="1"))
run(8192, 1, {"PFFFT_B200_CLUSTER_8192": "1"}); run(65536, 1, {"PFFFT_B200_CLUSTER": "0", "PFFFT_B200_TILED2D": "0"}); run(65536, 1, {}); run(32768, 1, {}); run(16384, 1, {"PFFFT_B200_TILED2D": "1"}); run(9216, 1, {}); run(36864, 1, {})
run(1024, 1, {}); run(4096, 0, {}); run(96, 0, {})
os.environ["PFFFT_B200_CONV_PIECE_KB"] = "16"
n, taps = 50000, 301
x = (np.arange(n) % 4093).astype(np.float32); y = np.zeros(n + 64, np.float32)
h = np.array([(-1.0, 1.0, 0.5)[j % 3] for j in range(taps)], np.float32)
fc = pf.FastConv(h, 1024, 0); got = fc.apply(x, y, n, 1); fc.close()
print("fastconv host pipeline produced=%d" % got, flush=True)
print("done")
