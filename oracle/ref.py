"""ctypes access to the checkers under oracle/ -- TEST INFRASTRUCTURE, never a product path.

Two libraries live here:
  oracle/_ref/libpffft_ref.so   the UNMODIFIED reference (marton78/pffft) compiled from
                                /root/reference/src/{pffft,pffft_double,pffft_common,pffastconv,fftpack}.c
                                by oracle/Makefile.  Built in the build container; travels to the
                                GPU box as a binary (git-ignored, not gpurun-ignored).
  oracle/liboracle.so           this repo's plain-C restatement (oracle/pffft_oracle.c).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libpffft_ref.so")
ORACLE_SO = os.path.join(HERE, "liboracle.so")

FORWARD, BACKWARD = 0, 1
REAL, COMPLEX = 0, 1


def have_ref():
    return os.path.exists(REF_SO)


def have_oracle():
    return os.path.exists(ORACLE_SO)


def _bind(lib, prefix, ctype):
    P = C.POINTER(ctype)
    g = lambda n: getattr(lib, prefix + n)
    g("new_setup").restype = C.c_void_p
    g("new_setup").argtypes = [C.c_int, C.c_int]
    g("destroy_setup").argtypes = [C.c_void_p]
    for n in ("transform", "transform_ordered"):
        g(n).argtypes = [C.c_void_p, P, P, P, C.c_int]
        g(n).restype = None
    g("zreorder").argtypes = [C.c_void_p, P, P, C.c_int]
    g("zreorder").restype = None
    for n in ("zconvolve_accumulate", "zconvolve_no_accu"):
        g(n).argtypes = [C.c_void_p, P, P, P, ctype]
        g(n).restype = None
    g("aligned_malloc").restype = C.c_void_p
    g("aligned_malloc").argtypes = [C.c_size_t]
    g("aligned_free").argtypes = [C.c_void_p]
    for n in ("min_fft_size",):
        g(n).argtypes = [C.c_int]
    g("is_valid_size").argtypes = [C.c_int, C.c_int]
    g("nearest_transform_size").argtypes = [C.c_int, C.c_int, C.c_int]
    g("next_power_of_two").argtypes = [C.c_int]
    g("is_power_of_two").argtypes = [C.c_int]
    g("simd_size").restype = C.c_int
    g("simd_arch").restype = C.c_char_p


def aligned(n, dtype):
    """numpy array on 64-byte aligned memory (the reference needs SIMD-aligned pointers, pffft.h:67-73);
    the returned view keeps its backing allocation alive through .base"""
    isz = np.dtype(dtype).itemsize
    raw = np.empty(n * isz + 64, dtype=np.uint8)
    off = (-raw.ctypes.data) % 64
    return raw[off:off + n * isz].view(dtype)


class PffftLib:
    """Thin object API over a pffft-ABI shared library (reference or restatement)."""

    def __init__(self, path):
        self.lib = C.CDLL(path)
        _bind(self.lib, "pffft_", C.c_float)
        self.has_double = hasattr(self.lib, "pffftd_new_setup")
        if self.has_double:
            _bind(self.lib, "pffftd_", C.c_double)
        if hasattr(self.lib, "pffastconv_new_setup"):
            L = self.lib
            L.pffastconv_new_setup.restype = C.c_void_p
            L.pffastconv_new_setup.argtypes = [C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int), C.c_int]
            L.pffastconv_destroy_setup.argtypes = [C.c_void_p]
            L.pffastconv_apply.restype = C.c_int
            L.pffastconv_apply.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float), C.c_int]

    def _pfx(self, dtype):
        return ("pffft_", C.c_float) if np.dtype(dtype) == np.float32 else ("pffftd_", C.c_double)

    def fn(self, dtype, name):
        return getattr(self.lib, self._pfx(dtype)[0] + name)

    # ---- transforms on one vector (numpy in, numpy out) -------------------------------------
    def new_setup(self, N, transform, dtype=np.float32):
        return self.fn(dtype, "new_setup")(N, transform)

    def destroy_setup(self, s, dtype=np.float32):
        self.fn(dtype, "destroy_setup")(s)

    def transform(self, N, transform, x, direction, ordered=True, dtype=np.float32):
        """x: 1-D array of N (real) or 2N (complex) scalars -> same-size output."""
        pfx, ct = self._pfx(dtype)
        n = N if transform == REAL else 2 * N
        s = self.new_setup(N, transform, dtype)
        if not s:
            raise ValueError("size rejected by pffft_new_setup: %d" % N)
        try:
            a = aligned(n, dtype); a[:] = np.asarray(x, dtype=dtype).ravel()
            o = aligned(n, dtype); w = aligned(n, dtype)
            P = C.POINTER(ct)
            f = self.fn(dtype, "transform_ordered" if ordered else "transform")
            f(s, a.ctypes.data_as(P), o.ctypes.data_as(P), w.ctypes.data_as(P), direction)
            return o.copy()
        finally:
            self.destroy_setup(s, dtype)

    def transform_batch(self, N, transform, x, direction, ordered=True, dtype=np.float32):
        """x: (batch, n) array; loops the single-vector call with one shared setup."""
        pfx, ct = self._pfx(dtype)
        n = N if transform == REAL else 2 * N
        x = np.asarray(x, dtype=dtype).reshape(-1, n)
        s = self.new_setup(N, transform, dtype)
        if not s:
            raise ValueError("size rejected by pffft_new_setup: %d" % N)
        out = np.empty_like(x)
        try:
            a = aligned(n, dtype); o = aligned(n, dtype); w = aligned(n, dtype)
            P = C.POINTER(ct)
            f = self.fn(dtype, "transform_ordered" if ordered else "transform")
            for b in range(x.shape[0]):
                a[:] = x[b]
                f(s, a.ctypes.data_as(P), o.ctypes.data_as(P), w.ctypes.data_as(P), direction)
                out[b] = o
            return out
        finally:
            self.destroy_setup(s, dtype)

    def zreorder(self, N, transform, x, direction, dtype=np.float32):
        pfx, ct = self._pfx(dtype)
        n = N if transform == REAL else 2 * N
        s = self.new_setup(N, transform, dtype)
        try:
            a = aligned(n, dtype); a[:] = np.asarray(x, dtype=dtype).ravel()
            o = aligned(n, dtype)
            P = C.POINTER(ct)
            self.fn(dtype, "zreorder")(s, a.ctypes.data_as(P), o.ctypes.data_as(P), direction)
            return o.copy()
        finally:
            self.destroy_setup(s, dtype)

    def zconvolve(self, N, transform, a_, b_, ab_, scaling, accumulate, dtype=np.float32):
        pfx, ct = self._pfx(dtype)
        n = N if transform == REAL else 2 * N
        s = self.new_setup(N, transform, dtype)
        try:
            a = aligned(n, dtype); a[:] = a_
            b = aligned(n, dtype); b[:] = b_
            ab = aligned(n, dtype); ab[:] = ab_
            P = C.POINTER(ct)
            f = self.fn(dtype, "zconvolve_accumulate" if accumulate else "zconvolve_no_accu")
            f(s, a.ctypes.data_as(P), b.ctypes.data_as(P), ab.ctypes.data_as(P), ct(scaling))
            return ab.copy()
        finally:
            self.destroy_setup(s, dtype)

    # ---- overlap-save convolution ------------------------------------------------------------
    def fastconv(self, h, x, block_len=0, flags=0, flush=1):
        """returns (y[:produced], produced, block_len_used); x holds len complex/real samples per flags."""
        h = np.ascontiguousarray(h, dtype=np.float32)
        x = np.ascontiguousarray(x, dtype=np.float32)
        cplx = bool(flags & 1)
        length = x.size // 2 if cplx else x.size
        bl = C.c_int(block_len)
        P = C.POINTER(C.c_float)
        s = self.lib.pffastconv_new_setup(h.ctypes.data_as(P), h.size, C.byref(bl), flags)
        if not s:
            return None, 0, bl.value
        try:
            y = np.full(x.size + 64, np.nan, dtype=np.float32)
            n = self.lib.pffastconv_apply(s, x.ctypes.data_as(P), length, y.ctypes.data_as(P), flush)
            return y[: n * (2 if cplx else 1)].copy(), n, bl.value
        finally:
            self.lib.pffastconv_destroy_setup(s)


_ref = None
_orc = None


def ref():
    """the unmodified reference library (kind 'reference')"""
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libpffft_ref.so missing: run `make -C oracle` where /root/reference exists")
        _ref = PffftLib(REF_SO)
    return _ref


def oracle():
    """this repo's C restatement (kind 'port')"""
    global _orc
    if _orc is None:
        if not have_oracle():
            raise RuntimeError("oracle/liboracle.so missing: run `make -C oracle`")
        _orc = PffftLib(ORACLE_SO)
    return _orc


def relmax(got, want):
    """parity metric of SURVEY 8(c): max|got-want| / max|want| (the reference validator's own, bench_pffft.c:372)"""
    got = np.asarray(got, dtype=np.float64); want = np.asarray(want, dtype=np.float64)
    d = np.max(np.abs(got - want)) if got.size else 0.0
    m = np.max(np.abs(want)) if want.size else 0.0
    return d / m if m > 0 else d
