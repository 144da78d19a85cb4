#!/usr/bin/env python
"""bench_configs.py -- secondary measurements: every BASELINE.json config (C1..C4) plus a size sweep,
one JSON line each (bench.py stays the single-line headline the driver reads).  Device-resident data,
CUDA-event timing, inputs larger than L2 where the config allows."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def ev_time(fn, iters, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-spectral", action="store_true")
    ap.add_argument("--only-spectral", action="store_true")
    args = ap.parse_args()
    import torch
    import pffft_b200 as pf
    peak = 6573.2
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    out = []

    def emit(d):
        print(json.dumps(d), flush=True)
        out.append(d)

    # spectral-domain kernels of the path: pffft_zreorder, pffft_zconvolve_accumulate/_no_accu,
    # batched, device resident, 1 GiB per array (>> L2).  Algorithmic bytes: zreorder reads 1 + writes 1 array;
    # zconvolve reads a (+ab when accumulating) and writes ab, b is ONE shared filter spectrum (L2 resident) or per-batch.
    def spectral_case(N, tr):
        per = N if tr == 0 else 2 * N
        batch = max(1, (1 << 30) // (4 * per))
        g = torch.Generator(device="cuda"); g.manual_seed(77)
        a = torch.rand((batch, per), generator=g, device="cuda") * 2 - 1
        b = torch.rand((batch, per), generator=g, device="cuda") * 2 - 1
        ab = torch.zeros_like(a)
        st = pf.Setup(N, tr)
        nb = batch * per * 4
        tag = "N=%d %s" % (N, "real" if tr == 0 else "cplx")
        for d, nm in ((0, "z->canonical"), (1, "canonical->z")):
            t = ev_time(lambda: pf.pffftb_zreorder_batch(st.handle, a, ab, batch, d), args.iters)
            emit({"config": "zreorder %s %s" % (nm, tag), "batch": batch, "ms": t * 1e3, "gbs_algorithmic": 2 * nb / t / 1e9,
                  "frac_of_hbm_peak": 2 * nb / t / 1e9 / peak})
        for acc, shared, arrays in ((1, 1, 3), (0, 1, 2), (1, 0, 4), (0, 0, 3)):
            bb = b[0].contiguous() if shared else b
            t = ev_time(lambda: pf.pffftb_zconvolve_batch(st.handle, a, bb, ab, 0.5, batch, shared, acc), args.iters)
            emit({"config": "zconvolve_%s %s b=%s" % ("accumulate" if acc else "no_accu", tag, "shared" if shared else "per-batch"),
                  "batch": batch, "ms": t * 1e3, "gbs_algorithmic": arrays * nb / t / 1e9,
                  "frac_of_hbm_peak": arrays * nb / t / 1e9 / peak})
        st.close()
        del a, b, ab
        torch.cuda.empty_cache()

    def dump():
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)

    if args.only_spectral:
        for N, tr in ((1024, 1), (4096, 0), (8192, 0)):
            spectral_case(N, tr)
        dump()
        return

    # C1: N=64 complex forward, batch 1, host pointers through the classic entry point (latency)
    x = (np.random.default_rng(1).random(128) * 2 - 1).astype(np.float32)
    y = np.empty_like(x)
    s = pf.pffft_new_setup(64, 1)
    for _ in range(20):
        pf.pffft_transform_ordered(s, x, y, None, 0)
    t0 = time.perf_counter()
    for _ in range(200):
        pf.pffft_transform_ordered(s, x, y, None, 0)
    dt = (time.perf_counter() - t0) / 200
    pf.pffft_destroy_setup(s)
    emit({"config": "C1 N=64 cplx fwd batch=1 host pointers (pffft_transform_ordered)", "us_per_call": dt * 1e6})

    def xform_case(name, N, tr, batch, direction, ordered, dtype=torch.float32):
        per = N if tr == 0 else 2 * N
        es = 4 if dtype == torch.float32 else 8
        g = torch.Generator(device="cuda"); g.manual_seed(1235)
        xin = torch.rand((batch, per), generator=g, device="cuda", dtype=dtype) * 2 - 1
        yout = torch.empty_like(xin)
        st = pf.Setup(N, tr, np.float32 if dtype == torch.float32 else np.float64)
        t = ev_time(lambda: pf.pffftb_transform_batch(st.handle, xin, yout, batch, direction, 1 if ordered else 0), args.iters)
        nbytes = 2 * batch * per * es
        flops = (5 if tr == 1 else 2.5) * N * np.log2(N) * batch
        emit({"config": name, "N": N, "transform": "real" if tr == 0 else "complex", "batch": batch, "kernel": st.kernel,
              "ms": t * 1e3, "ffts_per_s": batch / t, "gbs_algorithmic": nbytes / t / 1e9, "frac_of_hbm_peak": nbytes / t / 1e9 / peak,
              "gflops": flops / t / 1e9, "dtype": "f32" if dtype == torch.float32 else "f64",
              "direction": "fwd" if direction == 0 else "bwd", "ordered": bool(ordered)})
        st.close()
        del xin, yout
        torch.cuda.empty_cache()

    xform_case("C2 N=1024 cplx fwd batch=2^20", 1024, 1, 1 << 20, 0, True)
    xform_case("C2 N=1024 cplx bwd batch=2^20", 1024, 1, 1 << 20, 1, True)
    xform_case("C3 N=4096 real fwd batch=2^18", 4096, 0, 1 << 18, 0, True)
    xform_case("C3' N=4096 real bwd batch=2^18", 4096, 0, 1 << 18, 1, True)

    if not args.no_spectral:
        spectral_case(1024, 1)
        spectral_case(4096, 0)
        spectral_case(8192, 0)

    # C4: pffastconv 2^24-sample real stream, 4097 taps, device resident, one apply(flush=1) call
    n, taps = 1 << 24, 4097
    xs = torch.from_numpy((np.arange(n) % 4093).astype(np.float32)).cuda()
    ys = torch.empty(n, device="cuda")
    h = np.array([(-1.0, 1.0, 0.5)[j % 3] for j in range(taps)], np.float32)
    fc = pf.FastConv(h, 0, 0)
    produced = fc.apply(xs, ys, n, 1)
    t = ev_time(lambda: fc.apply(xs, ys, n, 1), args.iters)
    emit({"config": "C4 pffastconv 2^24 samples, 4097 taps, device resident", "Nfft": fc.block_len, "produced": produced,
          "ms": t * 1e3, "msamples_per_s": produced / t / 1e6, "gbs_algorithmic": 8.0 * produced / t / 1e9,
          "frac_of_hbm_peak": 8.0 * produced / t / 1e9 / peak})
    xh = xs.cpu().numpy(); yh = np.empty(n, np.float32)
    fc.apply(xh, yh, n, 1)
    t0 = time.perf_counter()
    for _ in range(3):
        fc.apply(xh, yh, n, 1)
    th = (time.perf_counter() - t0) / 3
    emit({"config": "C4 pffastconv, host pointers (pageable numpy buffers)", "ms": th * 1e3, "msamples_per_s": produced / th / 1e6})
    # same call with page-locked buffers from pffastconv_malloc (the allocator the reference API offers, pffastconv.h:176)
    import ctypes as C
    pf.lib.pffastconv_malloc.restype = C.c_void_p
    pf.lib.pffastconv_malloc.argtypes = [C.c_size_t]
    pf.lib.pffastconv_free.argtypes = [C.c_void_p]
    px, py = pf.lib.pffastconv_malloc(4 * n), pf.lib.pffastconv_malloc(4 * n)
    xp = np.ctypeslib.as_array(C.cast(px, C.POINTER(C.c_float)), shape=(n,))
    yp = np.ctypeslib.as_array(C.cast(py, C.POINTER(C.c_float)), shape=(n,))
    xp[:] = xh
    fc.apply(xp, yp, n, 1)
    t0 = time.perf_counter()
    for _ in range(5):
        fc.apply(xp, yp, n, 1)
    tp = (time.perf_counter() - t0) / 5
    same = bool(np.array_equal(yp[:produced], yh[:produced]))
    emit({"config": "C4 pffastconv, host pointers (page-locked buffers from pffastconv_malloc, pipelined H2D/kernel/D2H)",
          "ms": tp * 1e3, "msamples_per_s": produced / tp / 1e6, "h2d_bytes": 4 * n, "d2h_bytes": 4 * produced,
          "bit_identical_to_pageable_call": same})
    pf.lib.pffastconv_free(px); pf.lib.pffastconv_free(py)
    fc.close()

    # ---- the reference's CPU path beside each config (oracle/_ref on the host cores, bounded samples)
    if not args.no_cpu:
        import ctypes as C
        from oracle import ref as R
        so = os.path.join(ROOT, "oracle", "libcpubench.so")
        if R.have_ref() and os.path.exists(so):
            cb = C.CDLL(so)
            cb.cpu_bench_transform.restype = C.c_double
            cb.cpu_bench_transform.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint]
            cb.cpu_bench_fastconv.restype = C.c_double
            cb.cpu_bench_fastconv.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
            cores = len(os.sched_getaffinity(0))
            path = R.REF_SO.encode()
            for name, N, tr, sample, fi in (("C2 N=1024 cplx fwd", 1024, 1, 1 << 16, 0), ("C3 N=4096 real fwd", 4096, 0, 1 << 15, 0)):
                for nt in (1, cores):
                    smp = sample if nt > 1 else sample // 8
                    tsec = cb.cpu_bench_transform(path, N, tr, smp, nt, 8, fi, 1, 1235)
                    emit({"config": "CPU reference " + name, "threads": nt, "sample_transforms": smp, "passes": 8,
                          "ffts_per_s": smp * 8 / tsec, "gbs_algorithmic": smp * 8 * (16 * N if tr else 8 * N) / tsec / 1e9,
                          "kind": "reference (oracle/_ref, -O3 -march=haswell)"})
            prod = C.c_int(0)
            tsec = cb.cpu_bench_fastconv(path, 1 << 24, 4097, 3, C.byref(prod))
            emit({"config": "CPU reference C4 pffastconv 2^24 samples, 4097 taps (1 thread: setup not shareable)",
                  "ms": tsec * 1e3, "msamples_per_s": prod.value / tsec / 1e6, "produced": prod.value})

    if args.sweep:
        # ~1 GiB per buffer so that L2 (126 MB) cannot hold the working set
        for N in (16, 32, 64, 96, 128, 160, 192, 256, 384, 480, 512, 640, 800, 960, 1024, 2048, 4000, 4096, 8192, 12000,
                  16384, 36864, 65536):
            xform_case("sweep cplx fwd", N, 1, max(1, (1 << 30) // (8 * N)), 0, True)
        for N in (64, 192, 256, 512, 960, 1024, 1920, 2048, 4096, 8192, 16384, 65536, 131072):
            xform_case("sweep real fwd", N, 0, max(1, (1 << 30) // (4 * N)), 0, True)
        for N in (256, 1024, 4096):
            xform_case("sweep real bwd", N, 0, max(1, (1 << 30) // (4 * N)), 1, True)
        for N, tr in ((1024, 1), (4096, 1), (4096, 0), (96, 1), (8192, 1)):
            xform_case("sweep f64 fwd", N, tr, max(1, (1 << 30) // (16 * N)), 0, True, torch.float64)
        for N, tr in ((256, 1), (1024, 1), (4096, 1), (4096, 0), (960, 1)):
            xform_case("sweep fwd z-domain (pffft_transform)", N, tr, max(1, (1 << 30) // (8 * N)), 0, False)
            xform_case("sweep bwd z-domain (pffft_transform)", N, tr, max(1, (1 << 30) // (8 * N)), 1, False)
    dump()


if __name__ == "__main__":
    main()
