"""world_size-2 checks of the multi-process logic on CPU (gloo): batch sharding, table broadcast protocol and the
max-over-ranks timing reduction that bench.py uses under torchrun.  No GPU, no CUDA library calls."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        # 1) contiguous batch sharding: rank g owns [g*B/G, (g+1)*B/G)  
        B = 1000003
        lo, hi = B * rank // world, B * (rank + 1) // world
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([hi - lo]))
        total = int(sum(int(s.item()) for s in sizes))
        # 2) table broadcast: every rank builds tables; rank 0's overwrite the others bit-for-bit
        tables = torch.from_numpy(np.random.default_rng(rank).random(2048).astype(np.float32))
        mine_before = tables.clone()
        dist.broadcast(tables, src=0)
        ref0 = torch.from_numpy(np.random.default_rng(0).random(2048).astype(np.float32))
        # 3) timing reduction: value = units processed by all ranks / max over ranks of the elapsed time
        t = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        q.put((rank, total, bool(torch.equal(tables, ref0)), bool(torch.equal(mine_before, ref0)), float(t.item())))
    finally:
        dist.destroy_process_group()


def test_sharding_broadcast_and_max_reduce_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, total, same_after, same_before, tmax in res:
        assert total == 1000003                       # shards tile the batch exactly
        assert same_after                             # all ranks hold rank 0's tables
        assert same_before == (rank == 0)
        assert tmax == 2.0                            # max over ranks


def test_bench_reference_arm_non_zero_ranks_do_no_work(monkeypatch):
    """under torchrun the reference arm runs on rank 0 only; other ranks exit 0 without work"""
    import importlib
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    monkeypatch.setenv("RANK", "1")
    class A: steps = 1; warmup = 1; gpus = 2
    assert bench.run_reference_arm(A) == 0


# ------------------------------------------------------------------------------------------------------------------
# sharded overlap-save convolution (pffft_b200/sharded.py): halo exchange + per-rank pffastconv_apply == one global call
def _conv_worker(rank, world, port, q, total_len, ntaps):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        from oracle import ref as R
        from pffft_b200.sharded import ShardedStreamConv
        checker = R.ref() if R.have_ref() else R.oracle()

        def cpu_conv(taps, buf, out, feed_len, block_len, flush):   # the CPU checker stands in for the GPU call (no GPU here)
            y, n, _ = checker.fastconv(taps, buf.numpy()[:feed_len], block_len, 0, flush)
            out[:n] = torch.from_numpy(y)
            return n

        x = (np.arange(total_len) % 4093).astype(np.float32)   # tests/test_pffastconv.c:538-569
        h = np.array([(-1.0, 1.0, 0.5)[j % 3] for j in range(ntaps)], np.float32)
        sc = ShardedStreamConv(h, total_len, rank, world, block_len=0, conv=cpu_conv)
        buf = sc.alloc(device="cpu")
        sc.local(buf)[:] = torch.from_numpy(x[sc.lo:sc.hi])
        sc.exchange_halo(buf)
        halo_ok = bool(np.array_equal(buf.numpy(), x[sc.lo:sc.lo + sc.feed_len]))
        out = torch.full((sc.feed_len + 8,), float("nan"))
        n = sc.apply(buf, out)
        # the overlapped form (whole own blocks while the message is in flight, tail after it) gives the same samples
        buf2 = sc.alloc(device="cpu")
        sc.local(buf2)[:] = torch.from_numpy(x[sc.lo:sc.hi])
        out2 = torch.full((sc.feed_len + 8,), float("nan"))
        n2 = sc.exchange_and_apply(buf2, out2)
        same = n2 == n and bool(torch.equal(out2[:n], out[:n])) and bool(torch.isnan(out2[n:]).all())
        q.put((rank, sc.lo, n, halo_ok and same, out[:n].numpy().copy(), bool(torch.isnan(out[n:]).all())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,total_len,ntaps", [(2, 50000, 301), (3, 40001, 129), (2, 20000, 4097)])
def test_sharded_stream_convolution_equals_one_global_call(world, total_len, ntaps):
    import torch.multiprocessing as mp
    from oracle import ref as R
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_conv_worker, args=(r, world, port, q, total_len, ntaps)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=180) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    checker = R.ref() if R.have_ref() else R.oracle()
    x = (np.arange(total_len) % 4093).astype(np.float32)
    h = np.array([(-1.0, 1.0, 0.5)[j % 3] for j in range(ntaps)], np.float32)
    want, nwant, _ = checker.fastconv(h, x, 0, 0, 1)
    assert nwant == total_len - ntaps + 1
    got = np.concatenate([r[4] for r in res])
    assert got.size == nwant                                   # shards tile the valid output range exactly
    pos = 0
    for rank, lo, n, halo_ok, y, guard_ok in res:
        assert lo == pos and halo_ok and guard_ok
        pos += n
    # same algebra, different block phase per rank: float rounding differs, values agree to relmax 2e-6 (north_star
    # gate: 1e-5).  (The reference's own soft limit (max-min)/1e5, tests/test_pffastconv.c:685, is below one float ulp
    # of these ~1.4e6-sized sums for 4097 taps, so it cannot be the criterion here.)
    assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max()
