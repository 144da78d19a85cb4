"""pffft_b200 -- Python mirror of the C-ABI of libpffft_b200.so (test and bench harness).

The product is the shared library (pffft_b200/libpffft_b200.so, built from pffft_b200/csrc by
`make -C pffft_b200/csrc` or __graft_entry__.build()); this module only binds it with ctypes,
using the reference's own names and argument order (include/pffft/pffft.h, pffastconv.h of
marton78/pffft) so tests read like the reference's C tests.  Arrays may be numpy arrays (host
pointers) or torch CUDA tensors (device pointers); nothing here computes anything, and there is
no CPU fallback: if the library is missing, importing this module raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PFFFT_B200_LIB: A/B builds of the same library for tuning runs (e.g. the scalar-arithmetic build); never a CPU path
LIB_PATH = os.environ.get("PFFFT_B200_LIB") or os.path.join(_HERE, "libpffft_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "pffft_b200: %s not found -- build it with `make -C pffft_b200/csrc -j` (needs nvcc, sm_100a). "
        "There is no CPU fallback." % LIB_PATH)

lib = C.CDLL(LIB_PATH)

PFFFT_FORWARD, PFFFT_BACKWARD = 0, 1
PFFFT_REAL, PFFFT_COMPLEX = 0, 1
PFFASTCONV_CPLX_INP_OUT = 1
PFFASTCONV_CPLX_FILTER = 2
PFFASTCONV_DIRECT_INP = 4
PFFASTCONV_DIRECT_OUT = 8
PFFASTCONV_CPLX_SINGLE_FFT = 16
PFFASTCONV_SYMMETRIC = 32
PFFASTCONV_CORRELATION = 64

# every symbol the headers under include/pffft declare (checked by tests/test_abi.py)
_HELPERS = ["simd_size", "simd_arch", "min_fft_size", "next_power_of_two", "is_power_of_two", "is_valid_size",
            "nearest_transform_size", "aligned_malloc", "aligned_free"]
_CORE = ["new_setup", "destroy_setup", "transform", "transform_ordered", "zreorder", "zconvolve_accumulate",
         "zconvolve_no_accu"]
EXPORTED_SYMBOLS = (
    ["pffft_" + n for n in _CORE + _HELPERS] + ["pffftd_" + n for n in _CORE + _HELPERS] +
    ["pffastconv_new_setup", "pffastconv_destroy_setup", "pffastconv_apply", "pffastconv_malloc", "pffastconv_free",
     "pffastconv_simd_size"] +
    ["pffftb_transform_batch", "pffftdb_transform_batch", "pffftb_zreorder_batch", "pffftdb_zreorder_batch",
     "pffftb_zconvolve_batch", "pffftdb_zconvolve_batch", "pffftb_floats_per_transform",
     "pffftdb_doubles_per_transform", "pffftb_setup_device", "pffftb_setup_kernel", "pffftdb_setup_kernel",
     "pffftb_set_stream", "pffftdb_set_stream", "pffastconvb_set_stream", "pffftb_setup_tables",
     "pffftdb_setup_tables", "pffftb_last_error", "pffftb_launch_count", "pffftb_device_synchronize",
     # streaming / partitioned convolution under the boundary
     "pffastconvb_push", "pffastconvb_flush", "pffastconvb_pending", "pffastconvb_reset",
     "pffastconvb_partitioned_new", "pffastconvb_partitioned_destroy", "pffastconvb_partitioned_apply",
     "pffastconvb_partitioned_partitions", "pffastconvb_partitioned_set_stream",
     # multi-GPU
     "pffftb_multi_new", "pffftb_multi_destroy", "pffftb_multi_ngpus", "pffftb_multi_setup",
     "pffftb_multi_broadcast_backend", "pffftb_multi_transform_batch", "pffftb_multi_transform_shards",
     "pffftb_multi_synchronize", "pffftb_nccl_unique_id", "pffftb_setup_broadcast_tables"])

_vp = C.c_void_p


def _proto(name, restype, argtypes):
    f = getattr(lib, name)
    f.restype = restype
    f.argtypes = argtypes
    return f


for _pfx, _b, _sc in (("pffft_", "pffftb_", C.c_float), ("pffftd_", "pffftdb_", C.c_double)):
    _proto(_pfx + "new_setup", _vp, [C.c_int, C.c_int])
    _proto(_pfx + "destroy_setup", None, [_vp])
    _proto(_pfx + "transform", None, [_vp, _vp, _vp, _vp, C.c_int])
    _proto(_pfx + "transform_ordered", None, [_vp, _vp, _vp, _vp, C.c_int])
    _proto(_pfx + "zreorder", None, [_vp, _vp, _vp, C.c_int])
    _proto(_pfx + "zconvolve_accumulate", None, [_vp, _vp, _vp, _vp, _sc])
    _proto(_pfx + "zconvolve_no_accu", None, [_vp, _vp, _vp, _vp, _sc])
    _proto(_pfx + "simd_size", C.c_int, [])
    _proto(_pfx + "simd_arch", C.c_char_p, [])
    _proto(_pfx + "min_fft_size", C.c_int, [C.c_int])
    _proto(_pfx + "next_power_of_two", C.c_int, [C.c_int])
    _proto(_pfx + "is_power_of_two", C.c_int, [C.c_int])
    _proto(_pfx + "is_valid_size", C.c_int, [C.c_int, C.c_int])
    _proto(_pfx + "nearest_transform_size", C.c_int, [C.c_int, C.c_int, C.c_int])
    _proto(_pfx + "aligned_malloc", _vp, [C.c_size_t])
    _proto(_pfx + "aligned_free", None, [_vp])
    _proto(_b + "transform_batch", C.c_int, [_vp, _vp, _vp, C.c_size_t, C.c_int, C.c_int])
    _proto(_b + "zreorder_batch", C.c_int, [_vp, _vp, _vp, C.c_size_t, C.c_int])
    _proto(_b + "zconvolve_batch", C.c_int, [_vp, _vp, _vp, _vp, _sc, C.c_size_t, C.c_int, C.c_int])
    _proto(_b + "setup_kernel", C.c_char_p, [_vp])
    _proto(_b + "set_stream", C.c_int, [_vp, _vp])
    _proto(_b + "setup_tables", C.c_int, [_vp, C.POINTER(_vp), C.POINTER(C.c_size_t)])
_proto("pffftb_floats_per_transform", C.c_size_t, [_vp])
_proto("pffftdb_doubles_per_transform", C.c_size_t, [_vp])
_proto("pffftb_setup_device", C.c_int, [_vp])
_proto("pffftb_last_error", C.c_char_p, [])
_proto("pffftb_launch_count", C.c_ulonglong, [])
_proto("pffftb_device_synchronize", C.c_int, [])
_proto("pffastconv_new_setup", _vp, [_vp, C.c_int, C.POINTER(C.c_int), C.c_int])
_proto("pffastconv_destroy_setup", None, [_vp])
_proto("pffastconv_apply", C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int])
_proto("pffastconv_malloc", _vp, [C.c_size_t])
_proto("pffastconv_free", None, [_vp])
_proto("pffastconv_simd_size", C.c_int, [])
_proto("pffastconvb_set_stream", C.c_int, [_vp, _vp])
_proto("pffastconvb_push", C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int])
_proto("pffastconvb_flush", C.c_int, [_vp, _vp, C.c_int])
_proto("pffastconvb_pending", C.c_int, [_vp])
_proto("pffastconvb_reset", None, [_vp])
_proto("pffastconvb_partitioned_new", _vp, [_vp, C.c_int, C.c_int])
_proto("pffastconvb_partitioned_destroy", None, [_vp])
_proto("pffastconvb_partitioned_apply", C.c_longlong, [_vp, _vp, C.c_longlong, _vp])
_proto("pffastconvb_partitioned_partitions", C.c_int, [_vp])
_proto("pffastconvb_partitioned_set_stream", C.c_int, [_vp, _vp])
_proto("pffftb_multi_new", _vp, [C.c_int, C.c_int, C.c_int])
_proto("pffftb_multi_destroy", None, [_vp])
_proto("pffftb_multi_ngpus", C.c_int, [_vp])
_proto("pffftb_multi_setup", _vp, [_vp, C.c_int])
_proto("pffftb_multi_broadcast_backend", C.c_char_p, [_vp])
_proto("pffftb_multi_transform_batch", C.c_int, [_vp, _vp, _vp, C.c_size_t, C.c_int, C.c_int])
_proto("pffftb_multi_transform_shards", C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(C.c_size_t), C.c_int, C.c_int])
_proto("pffftb_multi_synchronize", C.c_int, [_vp])
_proto("pffftb_nccl_unique_id", C.c_int, [_vp])
_proto("pffftb_setup_broadcast_tables", C.c_int, [_vp, _vp, C.c_int, C.c_int])


def ptr(a):
    """address of a numpy array (host) or torch tensor (host or CUDA)"""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    if isinstance(a, int):
        return a
    raise TypeError("expected numpy array, torch tensor or address, got %r" % type(a))


def last_error():
    return lib.pffftb_last_error().decode()


def launch_count():
    return int(lib.pffftb_launch_count())


def _check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed: rc=%d %s" % (what, rc, last_error()))


def _pfx(dtype):
    dt = np.dtype(dtype)
    if dt == np.float32:
        return "pffft_", "pffftb_"
    if dt == np.float64:
        return "pffftd_", "pffftdb_"
    raise TypeError("pffft supports float32 and float64 only")


def _dtype_of(a):
    if isinstance(a, np.ndarray):
        return a.dtype
    import torch
    return {torch.float32: np.dtype(np.float32), torch.float64: np.dtype(np.float64)}[a.dtype]


# ------------------------------------------------------------------------------------------------
# function-style mirror of the reference API (same names, same argument order)
# ------------------------------------------------------------------------------------------------
def pffft_new_setup(N, transform, dtype=np.float32):
    """ref include/pffft/pffft.h:124 -- returns an opaque handle or None (size rejected / no device)"""
    return getattr(lib, _pfx(dtype)[0] + "new_setup")(int(N), int(transform))


def pffft_destroy_setup(setup, dtype=np.float32):
    getattr(lib, _pfx(dtype)[0] + "destroy_setup")(setup)


def pffft_transform(setup, inp, out, work, direction):
    """ref pffft.h:157 (z-domain result)"""
    getattr(lib, _pfx(_dtype_of(inp))[0] + "transform")(setup, ptr(inp), ptr(out), ptr(work), int(direction))


def pffft_transform_ordered(setup, inp, out, work, direction):
    """ref pffft.h:166 (canonical result)"""
    getattr(lib, _pfx(_dtype_of(inp))[0] + "transform_ordered")(setup, ptr(inp), ptr(out), ptr(work), int(direction))


def pffft_zreorder(setup, inp, out, direction):
    """ref pffft.h:180"""
    getattr(lib, _pfx(_dtype_of(inp))[0] + "zreorder")(setup, ptr(inp), ptr(out), int(direction))


def pffft_zconvolve_accumulate(setup, a, b, ab, scaling):
    """ref pffft.h:195"""
    getattr(lib, _pfx(_dtype_of(a))[0] + "zconvolve_accumulate")(setup, ptr(a), ptr(b), ptr(ab), scaling)


def pffft_zconvolve_no_accu(setup, a, b, ab, scaling):
    """ref pffft.h:209"""
    getattr(lib, _pfx(_dtype_of(a))[0] + "zconvolve_no_accu")(setup, ptr(a), ptr(b), ptr(ab), scaling)


def pffft_simd_size():
    return lib.pffft_simd_size()


def pffft_simd_arch():
    return lib.pffft_simd_arch().decode()


def pffft_min_fft_size(transform):
    return lib.pffft_min_fft_size(int(transform))


def pffft_next_power_of_two(N):
    return lib.pffft_next_power_of_two(int(N))


def pffft_is_power_of_two(N):
    return lib.pffft_is_power_of_two(int(N))


def pffft_is_valid_size(N, transform):
    return lib.pffft_is_valid_size(int(N), int(transform))


def pffft_nearest_transform_size(N, transform, higher):
    return lib.pffft_nearest_transform_size(int(N), int(transform), int(higher))


# batched extension (include/pffft/pffft_b200.h)
def pffftb_transform_batch(setup, inp, out, batch, direction, ordered=1):
    f = getattr(lib, _pfx(_dtype_of(inp))[1] + "transform_batch")
    _check(f(setup, ptr(inp), ptr(out), int(batch), int(direction), int(ordered)), "transform_batch")


def pffftb_zreorder_batch(setup, inp, out, batch, direction):
    f = getattr(lib, _pfx(_dtype_of(inp))[1] + "zreorder_batch")
    _check(f(setup, ptr(inp), ptr(out), int(batch), int(direction)), "zreorder_batch")


def pffftb_zconvolve_batch(setup, a, b, ab, scaling, batch, b_is_shared=0, accumulate=0):
    f = getattr(lib, _pfx(_dtype_of(a))[1] + "zconvolve_batch")
    _check(f(setup, ptr(a), ptr(b), ptr(ab), scaling, int(batch), int(b_is_shared), int(accumulate)), "zconvolve_batch")


def pffftb_setup_kernel(setup, dtype=np.float32):
    return getattr(lib, _pfx(dtype)[1] + "setup_kernel")(setup).decode()


def pffftb_setup_tables(setup, dtype=np.float32):
    """(device address, nbytes) of the plan's twiddle tables -- the buffer rank 0 broadcasts over NCCL"""
    p, n = _vp(), C.c_size_t()
    _check(getattr(lib, _pfx(dtype)[1] + "setup_tables")(setup, C.byref(p), C.byref(n)), "setup_tables")
    return p.value, n.value


def device_synchronize():
    _check(lib.pffftb_device_synchronize(), "device_synchronize")


# ------------------------------------------------------------------------------------------------
# small object wrappers (lifetime management only)
# ------------------------------------------------------------------------------------------------
class Setup:
    """owns a PFFFT_Setup / PFFFTD_Setup"""

    def __init__(self, N, transform, dtype=np.float32):
        self.N, self.transform, self.dtype = int(N), int(transform), np.dtype(dtype)
        self.handle = pffft_new_setup(N, transform, dtype)
        if not self.handle:
            raise ValueError("pffft_new_setup(%d, %s) returned NULL %s" % (
                N, "REAL" if transform == PFFFT_REAL else "COMPLEX", last_error()))
        self.per = self.N if transform == PFFFT_REAL else 2 * self.N

    def close(self):
        if getattr(self, "handle", None):
            try:
                pffft_destroy_setup(self.handle, self.dtype)
            except Exception:              # interpreter shutdown: module globals may already be gone
                pass
            self.handle = None

    def __del__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def kernel(self):
        return pffftb_setup_kernel(self.handle, self.dtype)

    def tables(self):
        return pffftb_setup_tables(self.handle, self.dtype)

    def _like(self, x):
        if isinstance(x, np.ndarray):
            return np.empty_like(x)
        import torch
        return torch.empty_like(x)

    def transform_batch(self, x, direction, ordered=True, out=None):
        """x: (batch, per) or (per,) numpy array / torch CUDA tensor -> same kind"""
        out = self._like(x) if out is None else out
        n = x.size if isinstance(x, np.ndarray) else x.numel()
        assert n % self.per == 0, "input size %d is not a multiple of %d" % (n, self.per)
        pffftb_transform_batch(self.handle, x, out, n // self.per, direction, 1 if ordered else 0)
        return out

    def zreorder_batch(self, x, direction, out=None):
        out = self._like(x) if out is None else out
        n = x.size if isinstance(x, np.ndarray) else x.numel()
        pffftb_zreorder_batch(self.handle, x, out, n // self.per, direction)
        return out

    def zconvolve_batch(self, a, b, ab, scaling, accumulate, b_is_shared=False):
        n = a.size if isinstance(a, np.ndarray) else a.numel()
        sc = float(scaling)
        pffftb_zconvolve_batch(self.handle, a, b, ab, sc, n // self.per, 1 if b_is_shared else 0, 1 if accumulate else 0)
        return ab


class FastConv:
    """owns a PFFASTCONV_Setup (ref include/pffft/pffastconv.h:145-173)"""

    def __init__(self, h, block_len=0, flags=0):
        h = np.ascontiguousarray(h, dtype=np.float32)
        bl = C.c_int(int(block_len))
        self.handle = lib.pffastconv_new_setup(h.ctypes.data, h.size, C.byref(bl), int(flags))
        self.block_len = bl.value
        self.flags = int(flags)
        self.filter_len = h.size

    def close(self):
        if getattr(self, "handle", None):
            try:
                lib.pffastconv_destroy_setup(self.handle)
            except Exception:
                pass
            self.handle = None

    def __del__(self):
        self.close()

    def apply(self, x, y, length, flush):
        """x, y: numpy arrays or torch CUDA tensors; length in (complex) samples; returns samples produced"""
        return lib.pffastconv_apply(self.handle, ptr(x), int(length), ptr(y), int(flush))

    # stateful stream (include/pffft/pffft_b200.h: pffastconvb_push / _flush)
    def push(self, x, length, y, capacity):
        n = lib.pffastconvb_push(self.handle, ptr(x), int(length), ptr(y), int(capacity))
        if n < 0:
            raise RuntimeError("pffastconvb_push failed: " + last_error())
        return n

    def flush(self, y, capacity):
        n = lib.pffastconvb_flush(self.handle, ptr(y), int(capacity))
        if n < 0:
            raise RuntimeError("pffastconvb_flush failed: " + last_error())
        return n

    @property
    def pending(self):
        return lib.pffastconvb_pending(self.handle)

    def reset(self):
        lib.pffastconvb_reset(self.handle)


class PartitionedConv:
    """owns a PFFASTCONVB_Partitioned (uniformly partitioned overlap-save, include/pffft/pffft_b200.h)"""

    def __init__(self, taps, part_len):
        taps = np.ascontiguousarray(taps, dtype=np.float32)
        self.filter_len = taps.size
        self.handle = lib.pffastconvb_partitioned_new(taps.ctypes.data, taps.size, int(part_len))
        if not self.handle:
            raise ValueError("pffastconvb_partitioned_new failed: " + last_error())
        self.partitions = lib.pffastconvb_partitioned_partitions(self.handle)

    def apply(self, x, y, length):
        n = lib.pffastconvb_partitioned_apply(self.handle, ptr(x), int(length), ptr(y))
        if n < 0:
            raise RuntimeError("pffastconvb_partitioned_apply failed: " + last_error())
        return int(n)

    def close(self):
        if getattr(self, "handle", None):
            try:
                lib.pffastconvb_partitioned_destroy(self.handle)
            except Exception:
                pass
            self.handle = None

    def __del__(self):
        self.close()


class Multi:
    """owns a PFFFTB_Multi: one plan per GPU of the node, tables broadcast once over NCCL (single process)"""

    def __init__(self, N, transform, ngpus=0):
        self.handle = lib.pffftb_multi_new(int(N), int(transform), int(ngpus))
        if not self.handle:
            raise ValueError("pffftb_multi_new failed: " + last_error())
        self.ngpus = lib.pffftb_multi_ngpus(self.handle)
        self.backend = lib.pffftb_multi_broadcast_backend(self.handle).decode()
        self.per = int(N) if transform == PFFFT_REAL else 2 * int(N)

    def setup(self, gpu):
        return lib.pffftb_multi_setup(self.handle, int(gpu))

    def transform_batch(self, x, out, batch, direction, ordered=1):
        _check(lib.pffftb_multi_transform_batch(self.handle, ptr(x), ptr(out), int(batch), int(direction), int(ordered)),
               "multi_transform_batch")

    def transform_shards(self, xs, outs, batches, direction, ordered=1):
        n = self.ngpus
        X = (_vp * n)(*[ptr(x) for x in xs]); O = (_vp * n)(*[ptr(o) for o in outs]); Bn = (C.c_size_t * n)(*[int(b) for b in batches])
        _check(lib.pffftb_multi_transform_shards(self.handle, X, O, Bn, int(direction), int(ordered)), "multi_transform_shards")

    def synchronize(self):
        _check(lib.pffftb_multi_synchronize(self.handle), "multi_synchronize")

    def close(self):
        if getattr(self, "handle", None):
            try:
                lib.pffftb_multi_destroy(self.handle)
            except Exception:
                pass
            self.handle = None

    def __del__(self):
        self.close()
