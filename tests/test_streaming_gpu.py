"""GPU tests of the streaming / partitioned convolution entry points of the C-ABI (pffastconvb_push / _flush,
pffastconvb_partitioned_*) against the UNMODIFIED reference (oracle/_ref): the reference's
pffastconv_apply over the whole stream, and the reference's pffft_zconvolve_accumulate as the partitioned inner loop."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def conv_limit(want):
    """the reference test's limit (max-min)/1e5 (tests/test_pffastconv.c:685), floored at a few float ulps of the largest
    output: with 4097 taps over the 4093-periodic ramp the outputs are ~1.4e6 with a range of only ~1e4, and the
    reference's formula would ask for 0.1 absolute where one float ulp is 0.125"""
    want = np.asarray(want, np.float64)
    return max((want.max() - want.min()) / 1e5, 8 * 2.0 ** -23 * np.abs(want).max())


def signal(n, taps):
    x = (np.arange(n) % 4093).astype(np.float32)                  # tests/test_pffastconv.c:538-569
    h = np.array([(-1.0, 1.0, 0.5)[j % 3] for j in range(taps)], np.float32)
    return x, h


@pytest.mark.parametrize("device", [False, True])
@pytest.mark.parametrize("taps,block_len,flags", [(129, 0, 0), (301, 2048, 0), (4097, 0, 0), (131, 0, 1), (131, 0, 17), (40, 64, 0)])
def test_stream_pushes_equal_the_reference_over_the_whole_stream(pf, ref, taps, block_len, flags, device):
    """any chunking + one flush: (1) bit-identical to ONE pffastconv_apply(flush=1) of this library over the whole stream,
    (2) equal to the reference's pffastconv_apply within its own test limit ((max-min)/1e5, tests/test_pffastconv.c:685).
    flags 1 = complex in/out (two real FFTs), 17 = complex single FFT."""
    import torch
    n = 100000 + 7
    w = 2 if flags & 1 else 1
    x, h = signal(n * w, taps)
    want_ref, n_ref, _ = ref.fastconv(h, x, block_len, flags, 1)
    fc = pf.FastConv(h, block_len, flags)
    one = np.full(n * w + 64, np.nan, np.float32)
    n_one = fc.apply(x, one, n, 1)
    assert n_one == n_ref
    rng = np.random.default_rng(taps)
    pos, outs = 0, []
    while pos < n:
        c = int(rng.integers(1, 30000))
        chunk = x[pos * w:(pos + c) * w]
        cn = chunk.size // w
        pos += cn
        cap = fc.pending + cn
        if device:
            y = torch.empty(max(cap * w, 1), device="cuda")
            got = fc.push(torch.from_numpy(chunk).cuda(), cn, y, cap)
            outs.append(y[:got * w].cpu().numpy())
        else:
            y = np.empty(max(cap * w, 1), np.float32)
            got = fc.push(chunk, cn, y, cap)
            outs.append(y[:got * w].copy())
    cap = fc.pending
    y = torch.empty(max(cap * w, 1), device="cuda") if device else np.empty(max(cap * w, 1), np.float32)
    got = fc.flush(y, cap)
    outs.append(y[:got * w].cpu().numpy() if device else y[:got * w].copy())
    assert fc.pending == (taps - 1 if n >= taps else n)
    # capacity check: a too small output buffer is an error, not a silent truncation
    fc.reset()
    with pytest.raises(RuntimeError):
        fc.push(x[:50000 * w], 50000, np.empty(16, np.float32), 8)
    fc.close()
    got_all = np.concatenate(outs)
    assert got_all.size == n_ref * w
    assert np.array_equal(got_all, one[:n_one * w])
    assert float(np.max(np.abs(got_all.astype(np.float64) - want_ref))) <= conv_limit(want_ref)


def _ref_partitioned(ref, x, h, B):
    """the partitioned scheme written with the REFERENCE's primitives: pffft_transform (z-domain), P calls of
    pffft_zconvolve_accumulate per block, pffft_transform backward"""
    F, N = h.size, 2 * B
    P = -(-F // B)
    n_out = x.size - F + 1
    K = -(-n_out // B)
    hr = np.zeros(P * B, np.float32); hr[:F] = h[::-1]
    xp = np.zeros((K + P) * B + N, np.float32); xp[:x.size] = x
    H = [None] * P
    for p in range(P):
        ht = np.zeros(N, np.float32)
        ht[(N - np.arange(B)) % N] = hr[p * B:(p + 1) * B]
        H[p] = ref.transform(N, 0, ht, 0, False)
    S = [ref.transform(N, 0, xp[k * B:k * B + N], 0, False) for k in range(K + P - 1)]
    y = np.empty(K * B, np.float32)
    for k in range(K):
        acc = np.zeros(N, np.float32)
        for p in range(P):
            acc = ref.zconvolve(N, 0, S[k + p], H[p], acc, 1.0 / N, True)
        y[k * B:(k + 1) * B] = ref.transform(N, 0, acc, 1, False)[:B]
    return y[:n_out]


@pytest.mark.parametrize("device", [True, False])
@pytest.mark.parametrize("taps,part", [(1000, 256), (64, 64), (300, 16), (2500, 512)])
def test_partitioned_convolution_vs_reference_primitives(pf, ref, taps, part, device):
    import torch
    n = 6000 + 11
    rng = np.random.default_rng(taps + part)
    x = (rng.random(n) * 2 - 1).astype(np.float32)
    h = (rng.random(taps) * 2 - 1).astype(np.float32)
    want = _ref_partitioned(ref, x, h, part)
    pc = pf.PartitionedConv(h, part)
    assert pc.partitions == -(-taps // part)
    if device:
        y = torch.full((n,), float("nan"), device="cuda")
        n0 = pf.launch_count()
        got_n = pc.apply(torch.from_numpy(x).cuda(), y, n)
        assert pf.launch_count() - n0 <= 4                          # forward + fused accumulate + backward (+ short last block)
        torch.cuda.synchronize()
        assert bool(torch.isnan(y[got_n:]).all())
        y = y[:got_n].cpu().numpy()
    else:
        yh = np.full(n, np.nan, np.float32)
        got_n = pc.apply(x, yh, n)
        y = yh[:got_n]
    pc.close()
    assert got_n == n - taps + 1 == want.size
    assert np.max(np.abs(y - want)) <= 1e-5 * np.max(np.abs(want))
    # direct sum in double at sampled positions (the reference's own check, tests/test_pffastconv.c:175-213)
    hr = h[::-1].astype(np.float64)
    for pos in (0, 1, part - 1, part, 2345, y.size - 1):
        w = float(np.dot(x[pos:pos + taps].astype(np.float64), hr))
        assert abs(y[pos] - w) <= 1e-5 * np.abs(want).max(), (pos, y[pos], w)


def test_partitioned_long_stream_vs_reference_pffastconv(pf, ref):
    """2^20 samples, 4097 taps in 9 partitions of 512: every output against the reference's single-FFT pffastconv_apply"""
    import torch
    n, taps, part = 1 << 20, 4097, 512
    rng = np.random.default_rng(5)
    x = (rng.random(n) * 2 - 1).astype(np.float32)
    h = (rng.random(taps) * 2 - 1).astype(np.float32)
    want, n_ref, _ = ref.fastconv(h, x, 0, 0, 1)
    pc = pf.PartitionedConv(h, part)
    y = torch.empty(n, device="cuda")
    got = pc.apply(torch.from_numpy(x).cuda(), y, n)
    pc.close()
    assert got == n_ref
    assert float(np.max(np.abs(y[:got].cpu().numpy() - want))) <= conv_limit(want)


def test_streaming_wrappers(pf):
    """the thin Python classes over the same entry points (pffft_b200/streaming.py)"""
    import torch
    from pffft_b200.streaming import PartitionedConv, StreamingConv
    x, h = signal(50000, 257)
    fc = pf.FastConv(h, 0, 0)
    want = np.empty(x.size, np.float32); nw = fc.apply(x, want, x.size, 1); fc.close()
    sc = StreamingConv(h)
    outs = [sc.push(torch.from_numpy(x[i:i + 7777]).cuda()).cpu().numpy() for i in range(0, x.size, 7777)]
    outs.append(sc.flush().cpu().numpy()); sc.close()
    assert np.array_equal(np.concatenate(outs), want[:nw])
    pc = PartitionedConv(h, 64)
    y = pc.apply(x)
    assert pc.launches <= 4 and y.size == nw and np.max(np.abs(y - want[:nw])) <= 1e-5 * np.abs(want[:nw]).max()
    pc.close()
