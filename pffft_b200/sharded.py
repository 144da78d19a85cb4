"""sharded.py -- pffastconv over a stream that is SHARDED across the GPUs of one box.

The reference convolves one host stream on one thread (`pffastconv_apply`, src/pffastconv.c:133-263) and leaves long
streams to the caller: "the caller re-feeds the unconsumed tail" (include/pffft/pffastconv.h:168-171).  With the stream
resident in the HBM of G GPUs -- rank r owns the contiguous samples [lo_r, hi_r) -- that tail is exactly the exchange
step of the path: output sample n needs x[n .. n+F-1], so rank r is short of the first F-1 samples of rank r+1.

    one halo message per rank pair (F-1 floats, e.g. 16 KiB for 4097 taps) sent with NCCL point-to-point over NVLink,
    then every rank runs the ordinary single-GPU `pffastconv_apply(flush=1)` on [own samples | halo].

Rank r then owns outputs [lo_r, hi_r) (the last rank stops at L-F+1): concatenated over ranks they are the samples a
single `pffastconv_apply(flush=1)` over the whole stream returns -- the same block algebra per rank, so values agree to
the rounding of a different block phase (tests compare against the reference at its own tolerance (max-min)/1e5,
tests/test_pffastconv.c:685).  No other inter-GPU traffic exists; the filter spectrum is built per rank from the same
taps (deterministic, so bit-identical everywhere).

This module is host-side orchestration only (torch.distributed for the halo message, the C-ABI for the compute).  The
`conv(taps, x, y, length, block_len, flush) -> produced` callable is injectable so the world_size-2 gloo test can run the
algebra on CPU with the oracle as the worker.
"""
import numpy as np


def shard_bounds(total_len, world):
    """contiguous, near-equal sample ranges: rank r owns [b[r], b[r+1])"""
    return [total_len * r // world for r in range(world + 1)]


class ShardedStreamConv:
    """Overlap-save FIR of a stream sharded over `world` ranks (real samples, PFFASTCONV flags = 0).

    rank r:  buf = alloc()                      # len_r + halo floats; fill buf[:len_r] with its samples
             exchange_halo(buf)                 # one send (to r-1) / one recv (from r+1)
             n = apply(buf, out)                # out[:n] = outputs [lo_r, lo_r + n)
         or  n = exchange_and_apply(buf, out)   # same samples, message overlapped with the blocks that do not need it
    """

    def __init__(self, taps, total_len, rank, world, block_len=0, group=None, conv=None):
        self.taps = np.ascontiguousarray(taps, dtype=np.float32)
        self.F = int(self.taps.size)
        self.total_len, self.rank, self.world, self.group = int(total_len), int(rank), int(world), group
        b = shard_bounds(self.total_len, self.world)
        self.lo, self.hi = b[rank], b[rank + 1]
        self.halo = self.F - 1
        if self.world > 1 and min(b[i + 1] - b[i] for i in range(self.world)) < self.halo:
            raise ValueError("every shard must hold at least filterLen-1 samples (halo comes from ONE neighbour)")
        # samples this rank feeds to pffastconv_apply: its own plus the halo (the last rank has no right neighbour)
        self.feed_len = (self.hi - self.lo) + (self.halo if rank + 1 < world else 0)
        # outputs it owns: n in [lo, hi) clipped to the global valid range [0, L-F+1)
        self.out_len = max(0, min(self.hi, self.total_len - self.F + 1) - self.lo)
        self.block_len = block_len
        self._conv = conv
        self._fc = None

    # ---- buffers -------------------------------------------------------------------------------------------------
    def alloc(self, device="cuda"):
        import torch
        return torch.zeros(self.feed_len, dtype=torch.float32, device=device)

    def local(self, buf):
        return buf[: self.hi - self.lo]

    # ---- the one exchange step of the path ------------------------------------------------------------------------
    def exchange_halo_begin(self, buf):
        """rank r+1 -> rank r: first F-1 samples.  NCCL send/recv (NVLink) on CUDA tensors, gloo on CPU tensors.
        Returns the pending requests; `exchange_halo_end` makes the current stream (or the host, on CPU) wait for them."""
        if self.world == 1 or self.halo == 0:
            return []
        import torch.distributed as dist
        ops = []
        if self.rank > 0:
            ops.append(dist.P2POp(dist.isend, buf[: self.halo].contiguous(), self.rank - 1, self.group))
        if self.rank + 1 < self.world:
            ops.append(dist.P2POp(dist.irecv, buf[self.hi - self.lo:], self.rank + 1, self.group))
        return dist.batch_isend_irecv(ops) if ops else []

    @staticmethod
    def exchange_halo_end(reqs):
        for w in reqs:
            w.wait()

    def exchange_halo(self, buf):
        self.exchange_halo_end(self.exchange_halo_begin(buf))

    # ---- compute: the ordinary single-GPU call ---------------------------------------------------------------------
    def _call(self, x, y, length, flush):
        if self._conv is not None:
            return self._conv(self.taps, x, y, length, self.block_len, flush)
        import pffft_b200 as pf
        if self._fc is None:
            self._fc = pf.FastConv(self.taps, self.block_len, 0)
            if not self._fc.handle:
                raise RuntimeError("pffastconv_new_setup failed: " + pf.last_error())
        return self._fc.apply(x, y, length, flush)

    def apply(self, buf, out):
        """out[:n] <- outputs [lo, lo+n) of the global convolution; returns n (== self.out_len).  Call after exchange_halo."""
        if self.feed_len < self.F:
            return 0
        n = self._call(buf, out, self.feed_len, 1)
        assert n == self.out_len, (n, self.out_len)
        return n

    def exchange_and_apply(self, buf, out):
        """the same result with the halo message HIDDEN behind the blocks that do not need it: every whole block inside
        this rank's own samples is convolved (applyFlush = 0) while the message is in flight; the remaining tail -- the
        only part that reads the halo -- follows with applyFlush = 1 at the block boundary the first call stopped at,
        i.e. with the same block phase, so the samples are those of `exchange_halo` + `apply` bit for bit."""
        own = self.hi - self.lo
        reqs = self.exchange_halo_begin(buf)
        n0 = self._call(buf, out, own, 0) if own >= self.F else 0
        self.exchange_halo_end(reqs)
        n1 = self._call(buf[n0:], out[n0:], self.feed_len - n0, 1) if self.feed_len - n0 >= self.F else 0
        assert n0 + n1 == self.out_len, (n0, n1, self.out_len)
        return n0 + n1

    def close(self):
        if self._fc is not None:
            self._fc.close()
            self._fc = None
