"""streaming.py -- the two callers either side of pffastconv that SURVEY 8f (row N4) lists after the hot path itself:

StreamingConv     the stateful form of `pffastconv_apply`'s contract.  The reference returns how many output samples a
                  call produced and expects the caller to feed the unconsumed input again, followed by new samples
                  (include/pffft/pffastconv.h:160-171; src/pffastconv.c:201, :262).  This class is that caller: it keeps
                  the unconsumed tail (at least filterLen-1 samples) between `push()` calls, so a stream of any chunking
                  yields exactly the samples of one call over the whole stream.

PartitionedConv   uniformly partitioned overlap-save for LONG filters: the taps are cut into P partitions of B taps, every
                  input window of 2B samples is transformed once, and each output block is
                        Y_k = sum_p  S_{k+p} * H_p          <- `pffft_zconvolve_accumulate` is this inner loop
                  in the z-domain (include/pffft/pffft.h:182-195 describes exactly this use), followed by one backward
                  transform per block.  Latency is B samples instead of the >= filterLen of the single-FFT scheme.
                  All spectra stay on the GPU; the P accumulate launches run over ALL blocks at once with the filter
                  partition as the shared operand.

Both are host-side orchestration over the C-ABI (batched transforms, batched zconvolve); no kernels live here."""
import numpy as np

import pffft_b200 as pf


def _is_torch(x):
    return not isinstance(x, np.ndarray)


class StreamingConv:
    """y = StreamingConv(taps).push(chunk) ... .flush(): real stream, same output convention as the reference:
    y[n] = sum_j x[n+j] * taps[F-1-j] (PFFASTCONV flags = 0), n counted over the whole stream."""

    def __init__(self, taps, block_len=0):
        self.fc = pf.FastConv(taps, block_len, 0)
        if not self.fc.handle:
            raise RuntimeError("pffastconv_new_setup failed: " + pf.last_error())
        self.F = int(np.asarray(taps).size)
        self.pending = None            # unconsumed input samples (same container kind as the chunks)
        self.consumed = 0              # stream position of pending[0] == number of outputs produced so far

    def _cat(self, a, b):
        if a is None:
            return b
        if _is_torch(b):
            import torch
            return torch.cat([a, b])
        return np.concatenate([a, b])

    def _run(self, flush):
        x = self.pending
        n_in = x.numel() if _is_torch(x) else x.size
        if n_in < self.F:
            return x[:0]
        if _is_torch(x):
            import torch
            x = x.contiguous()
            y = torch.empty(n_in, dtype=torch.float32, device=x.device)
        else:
            x = np.ascontiguousarray(x, dtype=np.float32)
            y = np.empty(n_in, np.float32)
        n = self.fc.apply(x, y, n_in, 1 if flush else 0)
        self.pending = x[n:]           # the caller "re-feeds the unconsumed tail" -- here, we are the caller
        self.consumed += n
        return y[:n]

    def push(self, chunk):
        """feed new samples (numpy array or CUDA tensor); returns the outputs that became available (whole blocks)"""
        self.pending = self._cat(self.pending, chunk)
        return self._run(False)

    def flush(self):
        """outputs for everything fed so far (applyFlush = 1); the last F-1 samples stay pending for later pushes"""
        if self.pending is None:
            return np.empty(0, np.float32)
        return self._run(True)

    def close(self):
        self.fc.close()


class PartitionedConv:
    """uniformly partitioned overlap-save convolution of a device-resident real stream; same output convention as
    pffastconv (y[n] = sum_j x[n+j] * taps[F-1-j], n in [0, L-F]) with partitions of `part_len` taps."""

    def __init__(self, taps, part_len):
        import torch
        taps = np.ascontiguousarray(taps, dtype=np.float32)
        B = int(part_len)
        if B < 16 or B & (B - 1):
            raise ValueError("part_len must be a power of two >= 16 (real transforms of 2*part_len >= 32 points)")
        self.F, self.B, self.N = int(taps.size), B, 2 * B
        self.P = (self.F + B - 1) // B
        self.setup = pf.Setup(self.N, pf.PFFFT_REAL)
        hr = np.zeros(self.P * B, np.float32)
        hr[: self.F] = taps[::-1]                                  # hr[j] = taps[F-1-j]
        ht = np.zeros((self.P, self.N), np.float32)
        idx = (self.N - np.arange(B)) % self.N                      # time-reversed, placed circularly (ref pffastconv.c:99-106)
        for p in range(self.P):
            ht[p, idx] = hr[p * B:(p + 1) * B]
        # partition spectra in the z-domain layout zconvolve works on (pffft_transform, not _ordered)
        self.H = self.setup.transform_batch(torch.from_numpy(ht).cuda(), pf.PFFFT_FORWARD, False)
        self.launches = 0

    def apply(self, x):
        """x: 1-D CUDA float tensor of L samples -> CUDA tensor of L-F+1 outputs"""
        import torch
        L = x.numel()
        n_out = L - self.F + 1
        if n_out <= 0:
            return x[:0]
        B, N, P = self.B, self.N, self.P
        K = (n_out + B - 1) // B                                    # output blocks
        Kw = K + P - 1                                              # input windows W_k = x[kB : kB + 2B]
        xp = torch.zeros((Kw + 1) * B, dtype=torch.float32, device=x.device)
        xp[: min(L, xp.numel())] = x[: min(L, xp.numel())]
        W = xp.unfold(0, N, B).contiguous()                          # (Kw, 2B) overlapping windows, materialised
        n0 = pf.launch_count()
        S = self.setup.transform_batch(W, pf.PFFFT_FORWARD, False)   # z-domain spectra of every window, one launch
        Y = torch.zeros((K, N), dtype=torch.float32, device=x.device)
        for p in range(P):                                           # Y_k += S_{k+p} * H_p / N   for all k at once
            self.setup.zconvolve_batch(S[p:p + K], self.H[p], Y, 1.0 / N, accumulate=True, b_is_shared=True)
        y = self.setup.transform_batch(Y, pf.PFFFT_BACKWARD, False)
        self.launches = pf.launch_count() - n0
        return y[:, :B].reshape(-1)[:n_out]

    def close(self):
        self.setup.close()
