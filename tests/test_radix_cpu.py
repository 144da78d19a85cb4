"""Compile-time-radix CTA kernels (pffft_b200/csrc/radix_kernels.cuh) stepped on the CPU through tests/emu: every core of
the table, every API mode, against the unmodified reference.  No GPU needed."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT, uniform

EMU = os.path.join(ROOT, "tests", "emu", "libemu.so")
CORES = [16, 48, 80, 144, 240, 400, 432, 720, 1152, 1200, 1280, 1296, 1440, 1600, 1728, 1920, 2000, 2160, 2304, 2400, 2560, 2592, 2880, 3200,
         3456, 3600, 3840, 4000, 4320, 4608, 4800, 5120, 5184, 5760, 6000, 6400, 6912, 7200, 7680, 8000, 9216, 9600, 10800, 11520, 12000, 12960, 13824, 14400]


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(EMU):
        pytest.skip("tests/emu/libemu.so not built (python __graft_entry__.py)")
    e = C.CDLL(EMU)
    e.emu_radix.argtypes = [C.c_int] * 4 + [C.c_void_p, C.c_void_p]
    return e


def _run(emu, N, tr, d, ordered, x):
    x = np.ascontiguousarray(x, np.float32)
    o = np.full_like(x, np.nan)
    assert emu.emu_radix(N, tr, d, ordered, x.ctypes.data, o.ctypes.data) == 0
    return o


@pytest.mark.parametrize("tr", [1, 0])
@pytest.mark.parametrize("core", CORES)
def test_every_core_and_mode_vs_reference(emu, ref, R, core, tr):
    N = core if tr == 1 else 2 * core
    rng = np.random.default_rng(core + tr)
    x = uniform(rng, 2 * core)
    fo = _run(emu, N, tr, 0, 1, x)
    fz = _run(emu, N, tr, 0, 0, x)
    assert R.relmax(fo, ref.transform(N, tr, x, 0, True)) <= 1e-5
    assert R.relmax(fz, ref.transform(N, tr, x, 0, False)) <= 1e-5
    assert np.array_equal(ref.zreorder(N, tr, fz, 0), fo)            # ordered == zreorder(unordered), bit-exact
    assert R.relmax(_run(emu, N, tr, 1, 1, fo), x * N) <= 1e-5
    assert R.relmax(_run(emu, N, tr, 1, 0, fz), x * N) <= 1e-5


# double-precision cores (pffft_b200/csrc/radix_d.cu): against numpy float64 at 1e-12 (the reference's own double path carries
# float-precision radix-3/5 constants)
CORES_D = [16, 32, 48, 64, 80, 96, 128, 144, 160, 192, 240, 256, 288, 320, 384, 400, 432, 480, 576, 640, 720, 768, 800, 864, 960, 1152, 1200,
           1280, 1296, 1440, 1600, 1728, 1920, 2000, 2160, 2304, 2400, 2560, 2592, 2880, 3456, 3600, 3840]


def _numpy_forward(x, N, tr):
    if tr == 1:
        W = np.fft.fft(x[0::2] + 1j * x[1::2])
        return np.stack([W.real, W.imag], -1).ravel()
    X = np.fft.rfft(x)
    w = np.stack([X.real[:-1], X.imag[:-1]], -1).ravel()
    w[1] = X.real[-1]
    return w


@pytest.mark.parametrize("tr", [1, 0])
@pytest.mark.parametrize("core", CORES_D)
def test_double_cores_vs_numpy(emu, ref, R, core, tr):
    emu.emu_radix_d.argtypes = [C.c_int] * 4 + [C.c_void_p, C.c_void_p]
    N = core if tr == 1 else 2 * core
    if tr == 0 and N % 32:
        pytest.skip("size")
    rng = np.random.default_rng(core + tr + 5)
    x = uniform(rng, 2 * core, np.float64)

    def run(d, ordered, v):
        v = np.ascontiguousarray(v, np.float64); o = np.full_like(v, np.nan)
        assert emu.emu_radix_d(N, tr, d, ordered, v.ctypes.data, o.ctypes.data) == 0
        return o
    fo = run(0, 1, x); fz = run(0, 0, x)
    assert R.relmax(fo, _numpy_forward(x, N, tr)) <= 1e-12
    assert np.array_equal(ref.zreorder(N, tr, fz, 0, np.float64), fo)
    assert R.relmax(run(1, 1, fo), x * N) <= 1e-12 and R.relmax(run(1, 0, fz), x * N) <= 1e-12
