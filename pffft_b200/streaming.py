"""streaming.py -- convenience wrappers over the C-ABI's streaming and partitioned convolution entry points
(include/pffft/pffft_b200.h: pffastconvb_push / _flush, pffastconvb_partitioned_*).

Everything that computes or keeps state lives in libpffft_b200.so (pffft_b200/csrc/fastconv.cu, partconv.cu) and is
reachable from C; these classes only allocate output arrays of the right kind (numpy for host pointers, torch CUDA
tensors for device pointers) around those calls."""
import numpy as np

import pffft_b200 as pf


def _is_torch(x):
    return not isinstance(x, np.ndarray)


def _empty_like_kind(x, n):
    if _is_torch(x):
        import torch
        return torch.empty(n, dtype=torch.float32, device=x.device)
    return np.empty(n, np.float32)


class StreamingConv:
    """y = StreamingConv(taps).push(chunk) ... .flush(): real stream, output convention of the reference:
    y[n] = sum_j x[n+j] * taps[F-1-j] (PFFASTCONV flags = 0), n counted over the whole stream.  The unconsumed tail of the
    stream is kept on the device by the library (pffastconvb_push), not here."""

    def __init__(self, taps, block_len=0):
        self.fc = pf.FastConv(taps, block_len, 0)
        if not self.fc.handle:
            raise RuntimeError("pffastconv_new_setup failed: " + pf.last_error())
        self.F = int(np.asarray(taps).size)
        self._kind = np.empty(0, np.float32)

    def push(self, chunk):
        """feed new samples (numpy array or CUDA tensor); returns the outputs of the blocks that became complete"""
        if not _is_torch(chunk):
            chunk = np.ascontiguousarray(chunk, dtype=np.float32)
        else:
            chunk = chunk.contiguous()
        self._kind = chunk[:0]
        n = chunk.numel() if _is_torch(chunk) else chunk.size
        cap = self.fc.pending + n
        y = _empty_like_kind(chunk, max(cap, 1))
        got = self.fc.push(chunk, n, y, cap)
        return y[:got]

    def flush(self):
        """outputs for everything fed so far (applyFlush = 1); the last F-1 samples stay pending for later pushes"""
        cap = self.fc.pending
        y = _empty_like_kind(self._kind, max(cap, 1))
        got = self.fc.flush(y, cap)
        return y[:got]

    def close(self):
        self.fc.close()


class PartitionedConv:
    """uniformly partitioned overlap-save convolution (pffastconvb_partitioned_*), same output convention as pffastconv"""

    def __init__(self, taps, part_len):
        self.pc = pf.PartitionedConv(taps, part_len)
        self.F = self.pc.filter_len
        self.P = self.pc.partitions
        self.launches = 0

    def apply(self, x):
        """x: 1-D numpy array or CUDA float tensor of L samples -> L-F+1 outputs of the same kind"""
        x = x.contiguous() if _is_torch(x) else np.ascontiguousarray(x, dtype=np.float32)
        L = x.numel() if _is_torch(x) else x.size
        n_out = max(L - self.F + 1, 0)
        y = _empty_like_kind(x, max(n_out, 1))
        n0 = pf.launch_count()
        got = self.pc.apply(x, y, L)
        self.launches = pf.launch_count() - n0
        return y[:got]

    def close(self):
        self.pc.close()
