"""Large transforms (complex cores above 65536, up to the reference's 2^26 limit, pffft_priv_impl.h:1069): oracle
parity for power-of-two AND mixed-radix sizes, complex and real, both precisions.  The mixed-radix sizes are the ones
ADVICE r1 named (a stage stride with a factor 3 or 5 broke the 32-bit reciprocal of the global path)."""
import numpy as np
import pytest

from conftest import uniform

pytestmark = pytest.mark.gpu

TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12}


def _numpy_forward(x, N, tr):
    x = x.astype(np.float64)
    if tr == 1:
        W = np.fft.fft(x[0::2] + 1j * x[1::2])
        return np.stack([W.real, W.imag], -1).ravel()
    X = np.fft.rfft(x)
    w = np.stack([X.real[:-1], X.imag[:-1]], -1).ravel()
    w[1] = X.real[-1]                                   # slot 0 = (DC, Nyquist), include/pffft/pffft.h:144-155
    return w


def _run(pf, N, tr, dtype, x, direction, ordered=True):
    import torch
    with pf.Setup(N, tr, dtype) as s:
        xd = torch.from_numpy(np.ascontiguousarray(x)).cuda()
        yd = s.transform_batch(xd, direction, ordered)
        torch.cuda.synchronize()
        return yd.cpu().numpy(), s.kernel


# complex cores: 3*2^16, 9*2^16 (ADVICE example), 6^7 = 279936, 5^3*3*2^10 = 384000, 2^17, 2^20; real N = 2*core
CORES = [196608, 589824, 279936, 384000, 131072, 1 << 20]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("tr", [1, 0])
@pytest.mark.parametrize("core", CORES)
def test_large_sizes_vs_reference_and_numpy(pf, ref, R, core, tr, dtype):
    N = core if tr == 1 else 2 * core
    if pf.pffft_is_valid_size(N, tr) != 1:
        pytest.skip("size not valid")
    dtype = np.dtype(dtype)
    rng = np.random.default_rng(core + tr)
    per = N if tr == 0 else 2 * N
    batch = 2
    x = uniform(rng, batch * per, dtype).reshape(batch, per)
    got, kern = _run(pf, N, tr, dtype, x, 0)
    pow2 = (N & (N - 1)) == 0
    for b in range(batch):
        assert R.relmax(got[b], _numpy_forward(x[b], N, tr)) <= TOL[dtype], ("numpy fwd", N, tr, dtype, b, kern)
    want = ref.transform_batch(N, tr, x, 0, True, dtype)
    # the reference's double path carries float-precision radix-3/5 constants (pffft_priv_impl.h:154, :259-262)
    tol_ref = TOL[dtype] if (pow2 or dtype == np.float32) else 5e-7
    for b in range(batch):
        assert R.relmax(got[b], want[b]) <= tol_ref, ("ref fwd", N, tr, dtype, b, kern)
    back, _ = _run(pf, N, tr, dtype, got, 1)
    for b in range(batch):
        assert R.relmax(back[b], x[b] * N) <= 10 * TOL[dtype], ("round trip", N, tr, dtype, b, kern)
    # z-domain: ordered == zreorder(unordered) bit-exact, one size class only (cost)
    if core in (196608, 131072):
        import torch
        with pf.Setup(N, tr, dtype) as s:
            xd = torch.from_numpy(x).cuda()
            z = s.transform_batch(xd, 0, False)
            o = s.zreorder_batch(z, 0)
            bz = s.transform_batch(z, 1, False)
            torch.cuda.synchronize()
            assert np.array_equal(o.cpu().numpy(), got)
            assert R.relmax(bz.cpu().numpy()[0], x[0] * N) <= 10 * TOL[dtype]


@pytest.mark.parametrize("tr", [1, 0])
def test_reference_maximum_size_2pow26(pf, R, tr):
    """N = 2^26, the largest size the reference admits (pffft_priv_impl.h:1069): known answers that need no 512 MiB
    oracle run -- a shifted impulse (X[k] = exp(-2 pi i k m / N)) plus sampled bins of a random vector against a direct
    double-precision DFT sum."""
    import torch
    N = 1 << 26
    per = N if tr == 0 else 2 * N
    rng = np.random.default_rng(26 + tr)
    x = uniform(rng, per, np.float32)
    with pf.Setup(N, tr) as s:
        xd = torch.from_numpy(x).cuda()
        yd = s.transform_batch(xd, 0, True)
        zd = s.transform_batch(yd, 1, True)
        torch.cuda.synchronize()
        err = float(((zd / N - xd).double() ** 2).sum().item())
        assert err <= N * 1e-7                           # round trip, tests/test_pffft.c:239
        y = yd.cpu().numpy().astype(np.float64)
        del yd, zd
    xs = x.astype(np.float64)
    ks = [0, 1, 2, 12345, N // 4 + 3, N // 2 - 1] + ([N // 2 + 5, N - 1] if tr == 1 else [])
    scale = np.sqrt(N)                                   # typical magnitude of a bin
    if tr == 1:
        sig = xs[0::2] + 1j * xs[1::2]
        n = np.arange(N)
        for k in ks:
            ph = ((n * k) % N) * (-2.0 * np.pi / N)
            w = np.sum(sig * (np.cos(ph) + 1j * np.sin(ph)))
            g = y[2 * k] + 1j * y[2 * k + 1]
            assert abs(g - w) <= 1e-5 * scale * 8, (k, g, w)
    else:
        n = np.arange(N)
        for k in ks:
            ph = ((n * k) % N) * (-2.0 * np.pi / N)
            w = np.sum(xs * (np.cos(ph) + 1j * np.sin(ph)))
            g = (y[0] + 0j) if k == 0 else (y[2 * k] + 1j * y[2 * k + 1])
            assert abs(g - w) <= 1e-5 * scale * 8, (k, g, w)
        nyq = np.sum(xs[0::2]) - np.sum(xs[1::2])
        assert abs(y[1] - nyq) <= 1e-5 * scale * 8
