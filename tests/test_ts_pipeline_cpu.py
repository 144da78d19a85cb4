"""Tiled Stockham pipeline (pffft_b200/csrc/ts_kernels.cuh) stepped on the CPU through tests/emu: the arithmetic of every
radix, the index algebra of first / later passes, the pre-/post-rotation stages, and the TICKET + DEPENDENCY-COUNTER protocol
under random interleavings (deadlock freedom, no read of an unfinished or recycled ring slot).  No GPU needed."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT, uniform

EMU = os.path.join(ROOT, "tests", "emu", "libemu.so")


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(EMU):
        pytest.skip("tests/emu/libemu.so not built (python __graft_entry__.py)")
    e = C.CDLL(EMU)
    e.emu_ts.argtypes = [C.c_int] * 5 + [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_uint]
    e.emu_ts_factorize.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    return e


def _factor(emu, Nc):
    P = C.c_int(0); A = (C.c_int * 4)()
    if not emu.emu_ts_factorize(Nc, C.byref(P), A):
        return None
    return [A[i] - 100 if A[i] >= 100 else 16 * A[i] for i in range(P.value)]


def test_factorisation_rules(emu, monkeypatch):
    monkeypatch.delenv("PFFFT_B200_TS_RADICES", raising=False)
    assert _factor(emu, 16384) == [128, 128]
    assert _factor(emu, 32768) == [256, 128]
    assert _factor(emu, 65536) == [256, 256]
    assert _factor(emu, 36864) == [192, 192]
    assert _factor(emu, 1 << 20) == [128, 128, 64]
    assert _factor(emu, 1 << 24) == [256, 256, 256]
    f = _factor(emu, 1 << 26)
    assert len(f) == 4 and np.prod(f) == 1 << 26
    assert _factor(emu, 589824) == [96, 96, 64]                 # 9 * 2^16
    assert _factor(emu, 250000) is None                         # 16 * 5^6: only one factor 16
    assert _factor(emu, 16 * 3 ** 8) is None
    for nc in (8192, 12288, 20480, 61440, 196608, 384000, 1 << 17, 1 << 22):
        f = _factor(emu, nc)
        assert f is not None and int(np.prod(f)) == nc and all(r <= 256 for r in f), (nc, f)
    assert _factor(emu, 384000) == [240, 160, 10]               # 2^10 * 3 * 5^3: two full passes + a closing radix 10


def _numpy_forward(x, N, tr):
    x = x.astype(np.float64)
    if tr == 1:
        W = np.fft.fft(x[0::2] + 1j * x[1::2])
        return np.stack([W.real, W.imag], -1).ravel()
    X = np.fft.rfft(x)
    w = np.stack([X.real[:-1], X.imag[:-1]], -1).ravel()
    w[1] = X.real[-1]
    return w


def _run(emu, prec, N, tr, d, ordered, x, batch, lag, window, seed=1):
    dt = np.float32 if prec == 0 else np.float64
    x = np.ascontiguousarray(x, dt)
    o = np.full_like(x, np.nan)
    rc = emu.emu_ts(prec, N, tr, d, ordered, x.ctypes.data, o.ctypes.data, batch, lag, window, seed)
    assert rc == 0, "emulated pipeline failed: rc=%d (-10 = deadlock)" % rc
    return o


# every radix 16*A of the library appears in one of these cores (first and later position)
RADIX_CASES = [("256,16", 4096), ("16,256", 4096), ("240,32", 7680), ("32,240", 7680), ("192,48", 9216), ("48,192", 9216),
               ("160,64", 10240), ("64,160", 10240), ("144,80", 11520), ("80,144", 11520), ("128,96", 12288), ("96,128", 12288),
               ("64,32,16", 32768), ("16,32,64", 32768), ("48,80,16", 61440), ("32,16,s10", 5120), ("16,16,s3", 768),
               ("16,32,s15", 7680), ("32,32,s9", 9216), ("16,16,s12", 3072), ("16,16,s2", 512), ("16,16,s4", 1024), ("16,16,s5", 1280),
               ("16,16,s6", 1536), ("16,16,s8", 2048)]


@pytest.mark.parametrize("radices,Nc", RADIX_CASES)
def test_every_radix_first_and_later(emu, R, monkeypatch, radices, Nc):
    monkeypatch.setenv("PFFFT_B200_TS_RADICES", radices)
    rng = np.random.default_rng(Nc)
    batch = 2
    x = uniform(rng, batch * 2 * Nc).reshape(batch, 2 * Nc)
    got = _run(emu, 0, Nc, 1, 0, 1, x, batch, 1, 8)
    for b in range(batch):
        assert R.relmax(got[b], _numpy_forward(x[b], Nc, 1)) <= 1e-5, (radices, b)
    back = _run(emu, 0, Nc, 1, 1, 1, got, batch, 1, 8)
    assert R.relmax(back, x * Nc) <= 1e-5


@pytest.mark.parametrize("tr", [1, 0])
@pytest.mark.parametrize("core", [16384, 36864])
def test_all_layouts_match_the_reference(emu, ref, R, monkeypatch, core, tr):
    """ordered and z-domain, forward and backward, complex and real: against the unmodified reference"""
    monkeypatch.delenv("PFFFT_B200_TS_RADICES", raising=False)
    N = core if tr == 1 else 2 * core
    rng = np.random.default_rng(core + tr)
    per = 2 * core
    x = uniform(rng, per)
    fo = _run(emu, 0, N, tr, 0, 1, x, 1, 2, 16)
    fz = _run(emu, 0, N, tr, 0, 0, x, 1, 2, 16)
    assert R.relmax(fo, ref.transform(N, tr, x, 0, True)) <= 1e-5
    wz = ref.transform(N, tr, x, 0, False)
    assert R.relmax(fz, wz) <= 1e-5
    assert np.array_equal(ref.zreorder(N, tr, fz, 0), fo)            # ordered == zreorder(unordered), bit-exact
    bo = _run(emu, 0, N, tr, 1, 1, fo, 1, 2, 16)
    bz = _run(emu, 0, N, tr, 1, 0, fz, 1, 2, 16)
    assert R.relmax(bo, x * N) <= 1e-5 and R.relmax(bz, x * N) <= 1e-5


def test_double_precision_1e12(emu, R, monkeypatch):
    monkeypatch.delenv("PFFFT_B200_TS_RADICES", raising=False)
    for tr, N in ((1, 20480), (0, 2 * 16384)):
        rng = np.random.default_rng(N)
        x = uniform(rng, 2 * N if tr == 1 else N, np.float64)
        got = _run(emu, 1, N, tr, 0, 1, x, 1, 1, 4)
        assert R.relmax(got, _numpy_forward(x, N, tr)) <= 1e-12


@pytest.mark.parametrize("lag,window", [(0, 1), (0, 40), (1, 3), (1, 64), (3, 200), (7, 24)])
def test_ticket_protocol_under_random_interleavings(emu, R, monkeypatch, lag, window):
    """many transforms through small rings (2*lag+1 slots, recycled many times) with `window` tickets in flight picked in
    random order: every output must still be right and some ticket must always be runnable"""
    monkeypatch.setenv("PFFFT_B200_TS_RADICES", "32,16,16")
    Nc, batch = 8192, 23
    rng = np.random.default_rng(lag * 100 + window)
    x = uniform(rng, batch * 2 * Nc).reshape(batch, 2 * Nc)
    want = np.stack([_numpy_forward(x[b], Nc, 1) for b in range(batch)])
    for seed in (1, 2, 3):
        got = _run(emu, 0, Nc, 1, 0, 1, x, batch, lag, window, seed)
        assert max(R.relmax(got[b], want[b]) for b in range(batch)) <= 1e-5
    # real forward (post-rotation stage reads the ring) and real backward (pre-rotation stage writes it)
    N = 2 * Nc
    xr = uniform(rng, 7 * N).reshape(7, N)
    fr = _run(emu, 0, N, 0, 0, 1, xr, 7, lag, window, 5)
    assert max(R.relmax(fr[b], _numpy_forward(xr[b], N, 0)) for b in range(7)) <= 1e-5
    br = _run(emu, 0, N, 0, 1, 1, fr, 7, lag, window, 6)
    assert R.relmax(br, xr * N) <= 1e-5
    # in place: input and output are the same buffer
    xi = x.copy()
    o = _run_inplace(emu, Nc, xi, batch, lag, window)
    assert max(R.relmax(o[b], want[b]) for b in range(batch)) <= 1e-5


def _run_inplace(emu, Nc, x, batch, lag, window):
    rc = emu.emu_ts(0, Nc, 1, 0, 1, x.ctypes.data, x.ctypes.data, batch, lag, window, 9)
    assert rc == 0
    return x


# ---------------------------------------------------------------------------------------------------------------
# warp-sized work items (pffft_b200/csrc/tsw_kernels.cuh): one warp = one tile, two phases around a __syncwarp
# ---------------------------------------------------------------------------------------------------------------
def _run_w(emu, N, tr, d, ordered, x, batch, lag, window, seed=1):
    emu.emu_tsw.argtypes = [C.c_int] * 4 + [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_uint]
    x = np.ascontiguousarray(x, np.float32)
    o = np.full_like(x, np.nan)
    rc = emu.emu_tsw(N, tr, d, ordered, x.ctypes.data, o.ctypes.data, batch, lag, window, seed)
    assert rc == 0, "emulated warp pipeline failed: rc=%d (-10 = deadlock, -2 = no warp plan)" % rc
    return o


# every radix (32, 64, 128, 256) as first, middle and last pass
TSW_CASES = [("256,256", 65536), ("128,128", 16384), ("256,128", 32768), ("128,256", 32768), ("64,32", 2048), ("32,64", 2048),
             ("64,64,32", 131072), ("32,64,128", 262144), ("256,32,64", 524288), ("32,32,32,32", 1 << 20)]


@pytest.mark.parametrize("radices,Nc", TSW_CASES)
def test_warp_items_every_radix_and_position(emu, R, monkeypatch, radices, Nc):
    monkeypatch.setenv("PFFFT_B200_TS_RADICES", radices)
    rng = np.random.default_rng(Nc + 7)
    batch = 2 if Nc <= 65536 else 1
    x = uniform(rng, batch * 2 * Nc).reshape(batch, 2 * Nc)
    got = _run_w(emu, Nc, 1, 0, 1, x, batch, 1, 24)
    for b in range(batch):
        assert R.relmax(got[b], _numpy_forward(x[b], Nc, 1)) <= 1e-5, (radices, b)
    back = _run_w(emu, Nc, 1, 1, 1, got, batch, 1, 24)
    assert R.relmax(back, x * Nc) <= 1e-5


@pytest.mark.parametrize("tr", [1, 0])
def test_warp_items_all_layouts_match_the_reference(emu, ref, R, monkeypatch, tr):
    monkeypatch.delenv("PFFFT_B200_TS_RADICES", raising=False)
    core = 16384
    N = core if tr == 1 else 2 * core
    rng = np.random.default_rng(core + tr + 100)
    x = uniform(rng, 2 * core)
    fo = _run_w(emu, N, tr, 0, 1, x, 1, 2, 64)
    fz = _run_w(emu, N, tr, 0, 0, x, 1, 2, 64)
    assert R.relmax(fo, ref.transform(N, tr, x, 0, True)) <= 1e-5
    assert R.relmax(fz, ref.transform(N, tr, x, 0, False)) <= 1e-5
    assert np.array_equal(ref.zreorder(N, tr, fz, 0), fo)
    bo = _run_w(emu, N, tr, 1, 1, fo, 1, 2, 64)
    bz = _run_w(emu, N, tr, 1, 0, fz, 1, 2, 64)
    assert R.relmax(bo, x * N) <= 1e-5 and R.relmax(bz, x * N) <= 1e-5


@pytest.mark.parametrize("lag,window", [(0, 1), (0, 90), (1, 7), (2, 300), (5, 64)])
def test_warp_items_protocol_under_random_interleavings(emu, R, monkeypatch, lag, window):
    monkeypatch.setenv("PFFFT_B200_TS_RADICES", "32,32,32")
    Nc, batch = 32768, 9
    rng = np.random.default_rng(lag * 100 + window + 3)
    x = uniform(rng, batch * 2 * Nc).reshape(batch, 2 * Nc)
    want = np.stack([_numpy_forward(x[b], Nc, 1) for b in range(batch)])
    for seed in (1, 2):
        got = _run_w(emu, Nc, 1, 0, 1, x, batch, lag, window, seed)
        assert max(R.relmax(got[b], want[b]) for b in range(batch)) <= 1e-5
