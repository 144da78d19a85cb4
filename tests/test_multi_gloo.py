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
