"""GPU tests of the tiled two-dimensional large-N plan (pffft_b200/csrc/tiled2d_kernels.cuh): default for complex cores
32768 and 65536, forced here for 16384 as well (PFFFT_B200_TILED2D=1).  Complex and real (the real wrappers run the
generic load / store passes around it), against the unmodified reference, round trip, in place, grid-stride loop."""
import os

import numpy as np
import pytest

from conftest import uniform

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tr", [1, 0])
@pytest.mark.parametrize("Nc", [16384, 32768, 65536])
def test_tiled2d_vs_reference(pf, ref, R, Nc, tr):
    import torch
    N = Nc if tr == 1 else 2 * Nc
    old = os.environ.get("PFFFT_B200_TILED2D")
    os.environ["PFFFT_B200_TILED2D"] = "1"
    try:
        s = pf.Setup(N, tr)
    finally:
        if old is None:
            os.environ.pop("PFFFT_B200_TILED2D", None)
        else:
            os.environ["PFFFT_B200_TILED2D"] = old
    try:
        assert s.kernel.startswith("tiled2d_"), s.kernel
        per = N if tr == 0 else 2 * N
        batch = 5
        x = uniform(np.random.default_rng(Nc + tr), batch * per).reshape(batch, per)
        xd = torch.from_numpy(x).cuda()
        y = s.transform_batch(xd, pf.PFFFT_FORWARD, True)
        z = s.transform_batch(y, pf.PFFFT_BACKWARD, True)
        torch.cuda.synchronize()
        want = ref.transform_batch(N, tr, x[:2], 0, True)
        for b in range(2):
            assert R.relmax(y[b].cpu().numpy(), want[b]) <= 1e-5
        assert float(((z / N - xd) ** 2).sum(dim=1).max()) <= N * 1e-7
        xi = xd.clone()
        s.transform_batch(xi, pf.PFFFT_FORWARD, True, out=xi)
        torch.cuda.synchronize()
        assert torch.equal(xi, y)                                   # in place == out of place
        big = max(8, (64 << 20) // (8 * Nc))                         # more tiles than resident CTAs: the grid-stride loop
        xb = xd[:1].repeat(big, 1).contiguous()
        yb = s.transform_batch(xb, pf.PFFFT_FORWARD, True)
        torch.cuda.synchronize()
        assert torch.equal(yb[-1], y[0]) and torch.equal(yb[big // 2], y[0])
    finally:
        s.close()


@pytest.mark.parametrize("Nc", [16384, 32768, 65536])
def test_tiled2d_cluster_fused_vs_reference(pf, ref, R, Nc):
    """PFFFT_B200_TILED2D=2: pass A hands its rows to pass C through DSMEM (8-CTA clusters).  First run on hardware at the start of round 2
    (profiles/r02_large_n.md): correct, 0.39-0.48 of the roofline -- not faster than the two-launch form, so it stays opt-in."""
    import torch
    old = os.environ.get("PFFFT_B200_TILED2D")
    os.environ["PFFFT_B200_TILED2D"] = "2"
    try:
        s = pf.Setup(Nc, 1)
    finally:
        if old is None:
            os.environ.pop("PFFFT_B200_TILED2D", None)
        else:
            os.environ["PFFFT_B200_TILED2D"] = old
    try:
        if not s.kernel.startswith("tiled2d_cluster8_"):
            pytest.skip("cluster shape not schedulable: " + s.kernel)
        batch = max(8, (64 << 20) // (8 * Nc)) + 3
        x = uniform(np.random.default_rng(Nc), 5 * 2 * Nc).reshape(5, 2 * Nc)
        xd = torch.from_numpy(x).cuda().repeat((batch + 4) // 5, 1)[:batch].contiguous()
        y = s.transform_batch(xd, pf.PFFFT_FORWARD, True)
        z = s.transform_batch(y, pf.PFFFT_BACKWARD, True)
        torch.cuda.synchronize()
        want = ref.transform_batch(Nc, 1, x[:2], 0, True)
        for b in range(2):
            assert R.relmax(y[b].cpu().numpy(), want[b]) <= 1e-5
        assert torch.equal(y[5:10], y[0:5]) and torch.equal(y[batch - batch % 5 - 5:batch - batch % 5], y[0:5])
        assert float(((z / Nc - xd) ** 2).sum(dim=1).max()) <= Nc * 1e-7
    finally:
        s.close()


@pytest.mark.parametrize("Nc", [7680, 9216, 12288, 20480, 24576, 36864, 40960, 49152, 61440, 16384, 65536])
def test_tiled2d_general_radix_vs_reference(pf, ref, R, Nc):
    """PFFFT_B200_TILED2D_GENERAL=1: Nc = 256*A1*A2 with radix-3/5 factors.  CPU-stepped only so far."""
    import torch
    old = os.environ.get("PFFFT_B200_TILED2D_GENERAL")
    os.environ["PFFFT_B200_TILED2D_GENERAL"] = "1"
    try:
        s = pf.Setup(Nc, 1)
    finally:
        if old is None:
            os.environ.pop("PFFFT_B200_TILED2D_GENERAL", None)
        else:
            os.environ["PFFFT_B200_TILED2D_GENERAL"] = old
    try:
        assert s.kernel.startswith("tiled2dg_"), s.kernel
        x = uniform(np.random.default_rng(Nc), 3 * 2 * Nc).reshape(3, 2 * Nc)
        xd = torch.from_numpy(x).cuda()
        y = s.transform_batch(xd, pf.PFFFT_FORWARD, True)
        z = s.transform_batch(y, pf.PFFFT_BACKWARD, True)
        torch.cuda.synchronize()
        want = ref.transform_batch(Nc, 1, x[:1], 0, True)
        assert R.relmax(y[0].cpu().numpy(), want[0]) <= 1e-5
        assert float(((z / Nc - xd) ** 2).sum(dim=1).max()) <= Nc * 1e-7
    finally:
        s.close()


@pytest.mark.parametrize("Nc", [16384, 32768, 65536])
def test_tiled2d_double_vs_numpy(pf, Nc):
    import torch
    old = os.environ.get("PFFFT_B200_TILED2D_GENERAL")
    os.environ["PFFFT_B200_TILED2D_GENERAL"] = "1"
    try:
        s = pf.Setup(Nc, 1, np.float64)
    finally:
        if old is None:
            os.environ.pop("PFFFT_B200_TILED2D_GENERAL", None)
        else:
            os.environ["PFFFT_B200_TILED2D_GENERAL"] = old
    try:
        assert s.kernel.startswith("tiled2dg_"), s.kernel
        x = uniform(np.random.default_rng(Nc), 2 * 2 * Nc, np.float64).reshape(2, 2 * Nc)
        y = s.transform_batch(torch.from_numpy(x).cuda(), pf.PFFFT_FORWARD, True).cpu().numpy()
        want = np.fft.fft(x[0].view(np.complex128))
        assert np.abs(y[0].view(np.complex128) - want).max() <= 1e-12 * np.abs(want).max()
    finally:
        s.close()
