"""CPU-side checks: the C-ABI library loads, exports every declared symbol, and its size algebra is
identical to the reference's (tests/test_fft_factors.c:36-61, tests/test_pffft.c:280-326)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import HAVE_GPU, ROOT


def test_library_loads_and_exports_every_declared_symbol(pf):
    for name in pf.EXPORTED_SYMBOLS:
        assert hasattr(pf.lib, name), "missing export " + name


def test_headers_and_exports_agree(pf):
    """every PFFFT_EXPORT / PFFASTCONV_EXPORT prototype in include/pffft/*.h is an exported symbol"""
    declared = set()
    for h in ("pffft.h", "pffft_double.h", "pffastconv.h", "pffft_b200.h"):
        txt = open(os.path.join(ROOT, "include", "pffft", h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        txt = re.sub(r"^\s*#.*$", "", txt, flags=re.M)          # drop the macro definitions themselves
        for m in re.finditer(r"(?:PFFFT_EXPORT|PFFASTCONV_EXPORT)\s+[^;(]*?\b(\w+)\s*\(", txt):
            declared.add(m.group(1))
    assert len(declared) >= 50
    for name in sorted(declared):
        assert hasattr(pf.lib, name), "header declares %s but the library does not export it" % name
    assert set(pf.EXPORTED_SYMBOLS) == declared


def test_enum_and_flag_values(pf):
    # ABI values fixed by the reference: include/pffft/pffft.h:108-117, pffastconv.h:83-134
    assert (pf.PFFFT_FORWARD, pf.PFFFT_BACKWARD, pf.PFFFT_REAL, pf.PFFFT_COMPLEX) == (0, 1, 0, 1)
    assert (pf.PFFASTCONV_CPLX_INP_OUT, pf.PFFASTCONV_CPLX_FILTER, pf.PFFASTCONV_DIRECT_INP, pf.PFFASTCONV_DIRECT_OUT,
            pf.PFFASTCONV_CPLX_SINGLE_FFT, pf.PFFASTCONV_SYMMETRIC, pf.PFFASTCONV_CORRELATION) == (1, 2, 4, 8, 16, 32, 64)


def test_simd_identity(pf):
    assert pf.pffft_simd_size() == 4          # layout granularity of the reference's SSE build
    assert pf.lib.pffftd_simd_size() == 4
    assert pf.lib.pffastconv_simd_size() == 4
    assert pf.pffft_simd_arch() == "sm_100a"
    assert pf.pffft_min_fft_size(pf.PFFFT_REAL) == 32 and pf.pffft_min_fft_size(pf.PFFFT_COMPLEX) == 16


def test_power_of_two_tables(pf):
    # tests/test_pffft.c:280-326
    ins = [1, 2, 3, 4, 5, 6, 7, 8, 9, 511, 512, 513]
    outs = [1, 2, 4, 4, 8, 8, 8, 8, 16, 512, 512, 1024]
    assert [pf.pffft_next_power_of_two(i) for i in ins] == outs
    assert [pf.lib.pffftd_next_power_of_two(i) for i in ins] == outs
    for i in range(0, 5000):
        assert pf.pffft_is_power_of_two(i) == (1 if i > 0 and (i & (i - 1)) == 0 else 0)


def test_size_algebra_matches_reference(pf, ref):
    for tr in (pf.PFFFT_REAL, pf.PFFFT_COMPLEX):
        nmin = pf.pffft_min_fft_size(tr)
        assert nmin == ref.lib.pffft_min_fft_size(tr)
        for N in list(range(0, 12 * nmin * 4 + 1)) + [4000, 4096, 12000, 36864, 65536, 1 << 20, 3 << 20, 5 << 22, 1 << 26]:
            assert pf.pffft_is_valid_size(N, tr) == ref.lib.pffft_is_valid_size(N, tr), (N, tr)
            assert pf.lib.pffftd_is_valid_size(N, tr) == ref.lib.pffftd_is_valid_size(N, tr), (N, tr)
        for N in list(range(1, 3000, 7)) + [100000, 1000001]:
            for hi in (0, 1):
                assert pf.pffft_nearest_transform_size(N, tr, hi) == ref.lib.pffft_nearest_transform_size(N, tr, hi)
    for N in list(range(-3, 70)) + [1 << 20, (1 << 20) + 1, (1 << 30) - 1]:
        assert pf.pffft_next_power_of_two(N) == ref.lib.pffft_next_power_of_two(N), N
        assert pf.pffft_is_power_of_two(N) == ref.lib.pffft_is_power_of_two(N), N


def test_aligned_malloc_contract(pf):
    # 64-byte aligned, free(NULL) safe (src/pffft_common.c:12-22)
    for nb in (1, 100, 4096, 1 << 20):
        p = pf.lib.pffft_aligned_malloc(nb)
        assert p and p % 64 == 0
        C.memset(p, 0xAB, nb)
        pf.lib.pffft_aligned_free(p)
    pf.lib.pffft_aligned_free(None)
    pf.lib.pffastconv_free(pf.lib.pffastconv_malloc(256))
    pf.lib.pffft_destroy_setup(None)
    pf.lib.pffftd_destroy_setup(None)
    pf.lib.pffastconv_destroy_setup(None)


def test_invalid_sizes_return_null(pf):
    # ref pffft_priv_impl.h:1066-1078, :1105-1109 -- rejected before any device work
    for N, tr in [(-16, 1), (0, 1), (8, 1), (24, 1), (16 * 7, 1), (16 * 11, 1), (48, 0), (32 * 7, 0), ((1 << 26) + 16, 1)]:
        assert not pf.pffft_new_setup(N, tr), (N, tr)
        assert not pf.pffft_new_setup(N, tr, np.float64), (N, tr)
    bl = C.c_int(0)
    h = np.ones(8, np.float32)
    assert not pf.lib.pffastconv_new_setup(h.ctypes.data, 8, C.byref(bl), pf.PFFASTCONV_CPLX_FILTER)   # ref pffastconv.c:71-72


@pytest.mark.skipif(HAVE_GPU, reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_a_gpu(pf):
    """The product has no CPU path: with no device a valid size yields NULL and an explanatory error."""
    s = pf.pffft_new_setup(1024, pf.PFFFT_COMPLEX)
    assert not s
    assert "CUDA" in pf.last_error() or "device" in pf.last_error()


@pytest.mark.parametrize("name", ["batch_c2c", "multi_gpu_c2c"])
def test_c_example_compiles_as_c99_against_the_headers(tmp_path, name):
    """examples/*.c: the batched and multi-GPU extensions are usable from plain C (prototypes only; running needs a GPU)"""
    import subprocess
    src = os.path.join(ROOT, "examples", name + ".c")
    obj = str(tmp_path / (name + ".o"))
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include", "pffft"), "-c", src, "-o", obj],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
