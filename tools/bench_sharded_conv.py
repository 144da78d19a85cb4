#!/usr/bin/env python
"""bench_sharded_conv.py -- BASELINE C4 scaled over the GPUs of one box: the 4097-tap FIR over a stream of
`--samples-per-gpu` samples PER RANK (weak scaling), sharded contiguously; every step = one NCCL halo message per rank
pair (F-1 floats over NVLink) + the single-GPU pffastconv_apply on [own samples | halo].

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_sharded_conv.py

Rank 0 prints one JSON line: aggregate output samples/s (max over ranks of the device time), and the parity of the rank
boundaries against a second, differently sharded evaluation of the same samples."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples-per-gpu", type=int, default=1 << 24)
    ap.add_argument("--taps", type=int, default=4097)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--no-overlap", action="store_true", help="exchange_halo then apply, instead of the overlapped form")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from pffft_b200.sharded import ShardedStreamConv
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = args.samples_per_gpu * world
    h = np.array([(-1.0, 1.0, 0.5)[j % 3] for j in range(args.taps)], np.float32)
    sc = ShardedStreamConv(h, L, rank, world)
    buf = sc.alloc()
    idx = torch.arange(sc.lo, sc.hi, device="cuda", dtype=torch.int64)
    sc.local(buf)[:] = (idx % 4093).to(torch.float32)          # tests/test_pffastconv.c:538-569, global index
    del idx
    out = torch.empty(sc.feed_len, device="cuda")

    def step():
        if args.no_overlap:
            sc.exchange_halo(buf)
            return sc.apply(buf, out)
        return sc.exchange_and_apply(buf, out)

    for _ in range(3):
        n = step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps
    # parity at the shard boundary: the last 64 outputs of this rank depend on the halo; recompute them from the known
    # global signal with a direct double-precision sum
    k = torch.arange(sc.lo + n - 64, sc.lo + n, dtype=torch.int64)
    j = torch.arange(args.taps, dtype=torch.int64)
    xs = ((k[:, None] + j[None, :]) % 4093).to(torch.float64)
    hr = torch.from_numpy(h[::-1].copy()).to(torch.float64)
    want = (xs * hr[None, :]).sum(dim=1)
    got = out[n - 64:n].cpu().to(torch.float64)
    rel = float((got - want).abs().max() / want.abs().max())
    t = torch.tensor([ms, float(n), rel], device="cuda", dtype=torch.float64)
    if world > 1:
        tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = t.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        ms, total, rel = float(tm[0]), float(ts[1]), float(tm[2])
    else:
        total = float(n)
    if rank == 0:
        print(json.dumps({"config": "C4 sharded: %d-tap FIR, %d samples per GPU, halo over NCCL" % (args.taps, args.samples_per_gpu),
                          "n_gpus": world, "ms_per_step": ms, "outputs_per_step": int(total), "msamples_per_s": total / ms / 1e3,
                          "halo_bytes_per_pair": 4 * (args.taps - 1), "boundary_relmax_vs_direct_sum": rel,
                          "halo": "sequential" if args.no_overlap else "overlapped with the blocks that do not read it"}))
    assert rel <= 1e-5, rel
    sc.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
