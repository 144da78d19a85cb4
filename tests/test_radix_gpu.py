"""Compile-time-radix CTA kernels on hardware (pffft_b200/csrc/radix_kernels.cuh): every core of the table, complex and
real, ordered and z-domain, forward and backward, batches that wrap the persistent grid, against the reference."""
import numpy as np
import pytest

from conftest import uniform

pytestmark = pytest.mark.gpu
CORES = [16, 48, 80, 144, 240, 400, 432, 720, 1152, 1200, 1280, 1296, 1440, 1600, 1728, 1920, 2000, 2160, 2304, 2400, 2560, 2592, 2880, 3200,
         3456, 3600, 3840, 4000, 4320, 4608, 4800, 5120, 5184, 5760, 6000, 6400, 6912, 7200, 7680, 8000, 9216, 9600, 10800, 11520, 12000, 12960, 13824, 14400]


@pytest.fixture()
def radix_on(monkeypatch):
    monkeypatch.setenv("PFFFT_B200_RADIX", "1")


@pytest.mark.parametrize("tr", [1, 0])
@pytest.mark.parametrize("core", CORES)
def test_every_core_and_mode_vs_reference(pf, ref, R, radix_on, core, tr):
    import torch
    N = core if tr == 1 else 2 * core
    rng = np.random.default_rng(core * 2 + tr)
    batch = max(5, min(40000, (64 << 20) // (8 * core)))              # ~64 MB: many times the resident CTAs
    x = uniform(rng, batch * 2 * core).reshape(batch, 2 * core)
    with pf.Setup(N, tr) as s:
        assert s.kernel.startswith("radix_"), s.kernel
        xd = torch.from_numpy(x).cuda()
        fo = s.transform_batch(xd, 0, True); fz = s.transform_batch(xd, 0, False)
        bo = s.transform_batch(fo, 1, True); bz = s.transform_batch(fz, 1, False)
        ro = s.zreorder_batch(fz, 0)
        inpl = xd.clone(); s.transform_batch(inpl, 0, True, out=inpl)
        torch.cuda.synchronize()
        assert torch.equal(ro, fo) and torch.equal(inpl, fo)
        assert float((bo / N - xd).abs().max()) <= 1e-5 and float((bz / N - xd).abs().max()) <= 1e-5
        idx = [0, 1, batch // 2, batch - 2, batch - 1]
        fo_, fz_ = fo[idx].cpu().numpy(), fz[idx].cpu().numpy()
    wo = ref.transform_batch(N, tr, x[idx], 0, True); wz = ref.transform_batch(N, tr, x[idx], 0, False)
    for j in range(len(idx)):
        assert R.relmax(fo_[j], wo[j]) <= 1e-5 and R.relmax(fz_[j], wz[j]) <= 1e-5


# double-precision cores (radix_d.cu): against numpy float64 at 1e-12 and by round trip
CORES_D = [16, 32, 48, 64, 80, 96, 128, 144, 160, 192, 240, 256, 288, 320, 384, 400, 432, 480, 576, 640, 720, 768, 800, 864, 960, 1152, 1200,
           1280, 1296, 1440, 1600, 1728, 1920, 2000, 2160, 2304, 2400, 2560, 2592, 2880, 3456, 3600, 3840]


def _numpy_forward(x, N, tr):
    if tr == 1:
        W = np.fft.fft(x[0::2] + 1j * x[1::2])
        return np.stack([W.real, W.imag], -1).ravel()
    X = np.fft.rfft(x)
    w = np.stack([X.real[:-1], X.imag[:-1]], -1).ravel()
    w[1] = X.real[-1]
    return w


@pytest.mark.parametrize("tr", [1, 0])
@pytest.mark.parametrize("core", CORES_D)
def test_double_cores_vs_numpy(pf, ref, R, core, tr):
    import torch
    N = core if tr == 1 else 2 * core
    if pf.pffft_is_valid_size(N, tr) != 1:
        pytest.skip("size")
    rng = np.random.default_rng(core * 3 + tr)
    batch = max(5, min(20000, (64 << 20) // (16 * core)))
    x = uniform(rng, batch * 2 * core, np.float64).reshape(batch, 2 * core)
    with pf.Setup(N, tr, np.float64) as s:
        assert s.kernel.startswith("radix_"), s.kernel
        xd = torch.from_numpy(x).cuda()
        fo = s.transform_batch(xd, 0, True); fz = s.transform_batch(xd, 0, False)
        bo = s.transform_batch(fo, 1, True); bz = s.transform_batch(fz, 1, False)
        ro = s.zreorder_batch(fz, 0)
        torch.cuda.synchronize()
        assert torch.equal(ro, fo)
        assert float((bo / N - xd).abs().max()) <= 1e-12 and float((bz / N - xd).abs().max()) <= 1e-12
        idx = [0, batch // 2, batch - 1]
        fo_ = fo[idx].cpu().numpy()
    for j, b in enumerate(idx):
        assert R.relmax(fo_[j], _numpy_forward(x[b], N, tr)) <= 1e-12
