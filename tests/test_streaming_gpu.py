"""GPU tests of the callers built on the path (pffft_b200/streaming.py): the stateful stream form of
pffastconv_apply and the uniformly partitioned convolution whose inner loop is pffft_zconvolve_accumulate."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def signal(n, taps):
    x = (np.arange(n) % 4093).astype(np.float32)                  # tests/test_pffastconv.c:538-569
    h = np.array([(-1.0, 1.0, 0.5)[j % 3] for j in range(taps)], np.float32)
    return x, h


@pytest.mark.parametrize("device", [False, True])
@pytest.mark.parametrize("taps,block_len", [(129, 0), (301, 2048), (4097, 0)])
def test_streaming_chunks_equal_one_call(pf, taps, block_len, device):
    """any chunking of the stream + one final flush == one pffastconv_apply(flush=1) over the whole stream, bit for bit
    (the class re-feeds the unconsumed tail exactly as pffastconv.h:160-171 asks the caller to)"""
    import torch
    from pffft_b200.streaming import StreamingConv
    n = 100000 + 7
    x, h = signal(n, taps)
    fc = pf.FastConv(h, block_len, 0)
    want = np.empty(n, np.float32)
    nw = fc.apply(x, want, n, 1)
    fc.close()
    assert nw == n - taps + 1
    sc = StreamingConv(h, block_len)
    rng = np.random.default_rng(taps)
    pos, outs = 0, []
    while pos < n:
        c = int(rng.integers(1, 30000))
        chunk = x[pos:pos + c]
        pos += chunk.size
        y = sc.push(torch.from_numpy(chunk).cuda() if device else chunk)
        outs.append(y.cpu().numpy() if device else np.array(y))
    y = sc.flush()
    outs.append(y.cpu().numpy() if device else np.array(y))
    sc.close()
    got = np.concatenate(outs)
    assert got.size == nw
    assert np.array_equal(got, want[:nw])


@pytest.mark.parametrize("taps,part", [(4097, 512), (1000, 256), (64, 64), (9000, 1024)])
def test_partitioned_convolution_vs_direct_sum_and_pffastconv(pf, taps, part):
    import torch
    from pffft_b200.streaming import PartitionedConv
    n = 60000 + 11
    rng = np.random.default_rng(taps + part)
    x = (rng.random(n) * 2 - 1).astype(np.float32)
    h = (rng.random(taps) * 2 - 1).astype(np.float32)
    pc = PartitionedConv(h, part)
    y = pc.apply(torch.from_numpy(x).cuda()).cpu().numpy()
    assert pc.P == -(-taps // part) and pc.launches == pc.P + 2     # 1 forward + P accumulate + 1 backward launch
    pc.close()
    assert y.size == n - taps + 1
    # direct sum in double at sampled positions (tests/test_pffastconv.c:175-213 is the reference's direct method)
    hr = h[::-1].astype(np.float64)
    for pos in (0, 1, part - 1, part, 12345, y.size - 1):
        want = float(np.dot(x[pos:pos + taps].astype(np.float64), hr))
        assert abs(y[pos] - want) <= 1e-5 * np.abs(y).max(), (pos, y[pos], want)
    # and the whole output against the single-FFT overlap-save of the same library
    fc = pf.FastConv(h, 0, 0)
    ref = np.empty(n, np.float32)
    nr = fc.apply(x, ref, n, 1)
    fc.close()
    assert nr == y.size
    assert np.abs(y - ref[:nr]).max() <= 1e-5 * np.abs(ref[:nr]).max()
