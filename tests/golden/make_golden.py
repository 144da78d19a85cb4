"""Generates tests/golden/pffft_golden.npz from the UNMODIFIED reference (oracle/_ref/libpffft_ref.so,
built from /root/reference by oracle/Makefile).  The reference ships no golden vectors of its own
(SURVEY F6), so these pin its behaviour for machines where /root/reference does not exist.

    python tests/golden/make_golden.py        # run in the build container

Inputs: uniform(-1,1) from numpy PCG64 with the seeds below (C1 uses seed 1); pffastconv uses
the reference test's ramp/pattern (tests/test_pffastconv.c:538-569)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref as R  # noqa: E402

CASES = [  # (N, transform, dtype, seed)
    (64, 1, "f4", 1), (16, 1, "f4", 2), (1024, 1, "f4", 1234), (4096, 0, "f4", 1235), (32, 0, "f4", 3),
    (96, 1, "f4", 4), (160, 0, "f4", 5), (4000, 1, "f4", 6), (2592, 0, "f4", 7), (8192, 0, "f4", 8),
    (1024, 1, "f8", 9), (4096, 0, "f8", 10), (64, 1, "f8", 11),
]


def main():
    r = R.ref()
    out = {}
    for N, tr, dt, seed in CASES:
        dtype = np.dtype(dt)
        n = N if tr == 0 else 2 * N
        rng = np.random.default_rng(seed)
        x = (rng.random(n) * 2 - 1).astype(dtype)
        key = "N%d_%s_%s" % (N, "r" if tr == 0 else "c", dt)
        fo = r.transform(N, tr, x, 0, True, dtype)
        fz = r.transform(N, tr, x, 0, False, dtype)
        out[key + "_x"] = x
        out[key + "_fwd_ordered"] = fo
        out[key + "_fwd_z"] = fz
        out[key + "_bwd_ordered"] = r.transform(N, tr, fo, 1, True, dtype)
    # zconvolve (accumulate and no_accu) on z-domain spectra, N=256 real and complex
    for tr in (0, 1):
        N = 256
        n = N if tr == 0 else 2 * N
        rng = np.random.default_rng(100 + tr)
        a, b, ab = [(rng.random(n) * 2 - 1).astype(np.float32) for _ in range(3)]
        key = "zconv_%s" % ("r" if tr == 0 else "c")
        out[key + "_a"], out[key + "_b"], out[key + "_ab"] = a, b, ab
        out[key + "_acc"] = r.zconvolve(N, tr, a, b, ab, 0.37, True)
        out[key + "_noacc"] = r.zconvolve(N, tr, a, b, ab, 0.37, False)
    # pffastconv: ramp input / (-1,1,0.5) taps like the reference test, three flag modes
    x = (np.arange(6000) % 4093).astype(np.float32)
    h = np.array([(-1.0, 1.0, 0.5)[j % 3] for j in range(131)], np.float32)
    for name, flags in (("real", 0), ("cplx2", 1), ("cplx1", 1 | 16)):
        for flush in (0, 1):
            y, n, bl = r.fastconv(h, x, 0, flags, flush)
            out["fc_%s_flush%d_y" % (name, flush)] = y
            out["fc_%s_flush%d_n" % (name, flush)] = np.array([n, bl])
    out["fc_x"], out["fc_h"] = x, h
    np.savez_compressed(os.path.join(HERE, "pffft_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
