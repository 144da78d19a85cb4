"""pffastconv_* parity (ref tests/test_pffastconv.c): returned lengths must equal the reference's (the hard
check there, :810-820), values within (max-min)/1e5 of the direct convolution (:685, :830-845)."""
import os

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def inputs(n, taps):
    # tests/test_pffastconv.c:538-569: X[i] = i % 4093, H = (-1, 1, 0.5, ...)
    x = (np.arange(n) % 4093).astype(np.float32)
    h = np.array([(-1.0, 1.0, 0.5)[j % 3] for j in range(taps)], np.float32)
    return x, h


def gpu_conv(pf, h, x, block_len, flags, flush, device_ptrs=False):
    cplx = bool(flags & pf.PFFASTCONV_CPLX_INP_OUT)
    length = x.size // 2 if cplx else x.size
    fc = pf.FastConv(h, block_len, flags)
    if not fc.handle:
        return None, 0, fc.block_len
    try:
        if device_ptrs:
            import torch
            xd = torch.from_numpy(x).cuda()
            yd = torch.full((x.size + 64,), float("nan"), device="cuda")
            n = fc.apply(xd, yd, length, flush)
            y = yd.cpu().numpy()
        else:
            y = np.full(x.size + 64, np.nan, np.float32)
            n = fc.apply(x, y, length, flush)
        k = n * (2 if cplx else 1)
        assert np.all(np.isnan(y[k:])), "wrote past the produced samples"      # NaN guard cells, test_pffastconv.c:118-131
        return y[:k].copy(), n, fc.block_len
    finally:
        fc.close()


def direct(x, h, n_out):
    # y[n] = sum_j x[n+j] * h[F-1-j]   == correlation with the reversed taps
    return np.correlate(x.astype(np.float64), h[::-1].astype(np.float64), mode="valid")[:n_out]


@pytest.mark.parametrize("flags_name", ["real", "cplx2", "cplx1"])
def test_lengths_and_values_vs_reference(pf, ref, flags_name):
    flags = {"real": 0, "cplx2": pf.PFFASTCONV_CPLX_INP_OUT,
             "cplx1": pf.PFFASTCONV_CPLX_INP_OUT | pf.PFFASTCONV_CPLX_SINGLE_FFT}[flags_name]
    n = 1 << 15
    for taps in range(124, 145, 4):                          # --quick range of the reference test (:915-936)
        x, h = inputs(n, taps)
        for block_len in (0, 512, 1024, 4096, 16384):        # block sizes 64..64K that fit (:490-500)
            for flush in (0, 1):
                want_y, want_n, want_bl = ref.fastconv(h, x, block_len, flags, flush)
                got_y, got_n, got_bl = gpu_conv(pf, h, x, block_len, flags, flush, device_ptrs=(taps % 8 == 0))
                assert got_bl == want_bl, (taps, block_len)
                assert got_n == want_n, (flags_name, taps, block_len, flush, got_n, want_n)
                if want_n == 0:
                    continue
                lim = (want_y.max() - want_y.min()) / 1e5
                assert np.max(np.abs(got_y - want_y)) <= lim, (flags_name, taps, block_len, flush)


def test_values_vs_direct_convolution(pf):
    x, h = inputs(20000, 131)
    y, n, bl = gpu_conv(pf, h, x, 0, 0, 1)
    assert n == 20000 - 131 + 1 and bl == 512     # 2*nextpow2(130)
    want = direct(x, h, n)
    assert np.max(np.abs(y - want)) <= (want.max() - want.min()) / 1e5
    # correlation flag: taps used as given (ref pffastconv.c:100-101)
    yc, nc, _ = gpu_conv(pf, h, x, 0, pf.PFFASTCONV_CORRELATION, 1)
    wantc = np.correlate(x.astype(np.float64), h.astype(np.float64), mode="valid")
    assert nc == n and np.max(np.abs(yc - wantc)) <= (wantc.max() - wantc.min()) / 1e5


def test_golden_fixture(pf):
    g = np.load(os.path.join(ROOT, "tests", "golden", "pffft_golden.npz"))
    x, h = g["fc_x"], g["fc_h"]
    for name, flags in (("real", 0), ("cplx2", 1), ("cplx1", 17)):
        for flush in (0, 1):
            want = g["fc_%s_flush%d_y" % (name, flush)]
            n_want, bl_want = g["fc_%s_flush%d_n" % (name, flush)]
            y, n, bl = gpu_conv(pf, h, x, 0, flags, flush)
            assert (n, bl) == (n_want, bl_want), (name, flush)
            if n:
                assert np.max(np.abs(y - want)) <= (want.max() - want.min()) / 1e5


def test_block_len_rounding_and_unsupported_flags(pf):
    import ctypes as C
    h = np.ones(100, np.float32)
    for req, want in ((0, 256), (100, 256), (300, 512), (512, 512), (513, 1024)):   # ref pffastconv.c:62-80
        fc = pf.FastConv(h, req, 0)
        assert fc.handle and fc.block_len == want
        fc.close()
    fc = pf.FastConv(np.ones(3, np.float32), 0, 0)
    assert fc.block_len == 32                                 # minimum FFT length
    fc.close()
    assert not pf.FastConv(h, 0, pf.PFFASTCONV_CPLX_FILTER).handle
    # DIRECT_INP/OUT are copy-elision hints: identical results (the reference itself crashes on DIRECT_INP)
    x, hh = inputs(5000, 64)
    y0, n0, _ = gpu_conv(pf, hh, x, 0, 0, 1)
    y1, n1, _ = gpu_conv(pf, hh, x, 0, pf.PFFASTCONV_DIRECT_INP | pf.PFFASTCONV_DIRECT_OUT, 1)
    assert n0 == n1 and np.array_equal(y0, y1)


def test_c4_config_sampled(pf):
    """BASELINE C4: 2^24-sample ramp, 4097 taps -> Nfft 8192, 16 773 120 outputs; sampled direct sums in double."""
    import torch
    n, taps = 1 << 24, 4097
    x, h = inputs(n, taps)
    fc = pf.FastConv(h, 0, 0)
    assert fc.block_len == 8192
    xd = torch.from_numpy(x).cuda()
    yd = torch.empty(n, device="cuda")
    produced = fc.apply(xd, yd, n, 1)
    assert produced == n - taps + 1 == 16773120
    y = yd[:produced].cpu().numpy()
    fc.close()
    rng = np.random.default_rng(0)
    pos = np.concatenate([[0, 1, 4095, 4096, produced - 1], rng.integers(0, produced, 40)])
    hr = h[::-1].astype(np.float64)
    # The reference test's limit (max-min)/1e5 (:685) is meant for its short filters; with 4097 taps on the
    # 0..4092 ramp the outputs sit near 1.4e6 (one float ulp = 0.125) while their range is only ~7e3, so that
    # limit is below float resolution for ANY implementation, the reference included.  Use the north-star
    # metric instead: |err| <= 1e-5 * max|y|  (observed: ~1e-7, i.e. 1 ulp).
    lim = 1e-5 * float(np.abs(y).max())
    worst = 0.0
    for p in pos:
        want = float(np.dot(x[p:p + taps].astype(np.float64), hr))
        worst = max(worst, abs(float(y[p]) - want))
    assert worst <= lim, (worst, lim)
    assert worst <= 4 * float(np.spacing(np.float32(np.abs(y).max()))), worst   # and in fact within a few ulp


@pytest.mark.parametrize("flags_name", ["real", "cplx_single"])
@pytest.mark.parametrize("flush", [0, 1])
def test_host_pointer_pipeline_equals_device_call(pf, flags_name, flush):
    """host-pointer calls cut the stream into pieces that overlap H2D / kernel / D2H on three streams; every piece size
    (one block per piece up to the whole stream) must give exactly the samples of the one-launch device-pointer call"""
    flags = 0 if flags_name == "real" else (pf.PFFASTCONV_CPLX_INP_OUT | pf.PFFASTCONV_CPLX_SINGLE_FFT)
    taps = 301
    n = 70000 + 37                                      # real samples / complex samples
    x, h = inputs(n * (2 if flags else 1), taps)
    want, n_dev, bl = gpu_conv(pf, h, x, 1024, flags, flush, device_ptrs=True)
    assert n_dev > 0
    old = os.environ.get("PFFFT_B200_CONV_PIECE_KB")
    try:
        for kb in ("1", "4", "24", "100", "100000"):    # 1 KiB -> one block per piece
            os.environ["PFFFT_B200_CONV_PIECE_KB"] = kb
            got, n_host, _ = gpu_conv(pf, h, x, 1024, flags, flush, device_ptrs=False)
            assert n_host == n_dev, (kb, n_host, n_dev)
            assert np.array_equal(got, want), kb
    finally:
        if old is None:
            os.environ.pop("PFFFT_B200_CONV_PIECE_KB", None)
        else:
            os.environ["PFFFT_B200_CONV_PIECE_KB"] = old
