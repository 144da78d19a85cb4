"""GPU tests of the thread-block-cluster plans (pffft_b200/csrc/cluster_kernels.cuh): complex cores 8192..65536 held in the
distributed shared memory of a cluster.  Every variant (cluster size, rows per CTA, DSMEM row scatter vs strided reads)
is forced through its environment switch and checked against the unmodified reference (relmax <= 1e-5, north_star),
bit-exactly against itself for batch position independence, and through the real / z-domain wrappers."""
import os

import numpy as np
import pytest

from conftest import uniform

pytestmark = pytest.mark.gpu

# (N complex, env overrides, expected kernel-name prefix)
VARIANTS = [
    (16384, {"PFFFT_B200_CLUSTER_MODE": "1"}, "cluster4_4x4096_dsmem_rows"),
    (16384, {"PFFFT_B200_CLUSTER_MODE": "0"}, "cluster4_4x4096"),
    (32768, {"PFFFT_B200_CLUSTER": "all", "PFFFT_B200_CLUSTER_MODE": "1"}, "cluster8_8x4096_dsmem_rows"),
    (32768, {"PFFFT_B200_CLUSTER": "all", "PFFFT_B200_CLUSTER_MODE": "0"}, "cluster8_8x4096"),
    (65536, {"PFFFT_B200_CLUSTER": "all"}, "cluster8_16x4096"),
    (32768, {"PFFFT_B200_CLUSTER": "all", "PFFFT_B200_CLUSTER_SHAPE": "4x2"}, "cluster4_8x4096"),
    (65536, {"PFFFT_B200_CLUSTER": "all", "PFFFT_B200_CLUSTER_SHAPE": "4x4"}, "cluster4_16x4096"),
    (65536, {"PFFFT_B200_CLUSTER": "all", "PFFFT_B200_CLUSTER_R16": "16", "PFFFT_B200_CLUSTER_MODE": "1"}, "cluster16_16x4096_dsmem_rows"),
    (65536, {"PFFFT_B200_CLUSTER": "all", "PFFFT_B200_CLUSTER_R16": "16", "PFFFT_B200_CLUSTER_MODE": "0"}, "cluster16_16x4096"),
    (8192, {"PFFFT_B200_CLUSTER_8192": "1", "PFFFT_B200_CLUSTER_MODE": "1"}, "cluster2_2x4096_dsmem_rows"),
    (8192, {"PFFFT_B200_CLUSTER_8192": "1", "PFFFT_B200_CLUSTER_MODE": "0"}, "cluster2_2x4096"),
]
ENV_KEYS = ("PFFFT_B200_CLUSTER", "PFFFT_B200_CLUSTER_MODE", "PFFFT_B200_CLUSTER_R16", "PFFFT_B200_CLUSTER_8192", "PFFFT_B200_CLUSTER_SHAPE",
            "PFFFT_B200_TILED2D")


class env_set:
    def __init__(self, kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in ENV_KEYS}
        for k in ENV_KEYS:
            os.environ.pop(k, None)
        os.environ["PFFFT_B200_TILED2D"] = "0"          # these tests address the cluster / split plans of the same sizes
        os.environ.update(self.kv)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def make_setup(pf, N, tr, env, want_name):
    """plan under the environment switches (read at plan creation); skips when the device cannot schedule the shape"""
    with env_set(env):
        s = pf.Setup(N, tr)
    if want_name is not None and not s.kernel.startswith("cluster"):
        s.close()
        pytest.skip("cluster shape not schedulable on this device (plan fell back to %s)" % s.kernel)
    return s


@pytest.mark.parametrize("N,env,name", VARIANTS)
def test_cluster_variants_vs_reference(pf, ref, R, N, env, name):
    import torch
    s = make_setup(pf, N, 1, env, name)
    try:
        assert s.kernel == name, s.kernel
        rng = np.random.default_rng(N + len(env))
        batch = 5                                           # not a multiple of anything
        x = uniform(rng, batch * 2 * N).reshape(batch, 2 * N)
        xd = torch.from_numpy(x).cuda()
        yd = s.transform_batch(xd, pf.PFFFT_FORWARD, True)
        zd = s.transform_batch(yd, pf.PFFFT_BACKWARD, True)
        torch.cuda.synchronize()
        y = yd.cpu().numpy()
        want = ref.transform_batch(N, 1, x[:2], 0, True)
        for b in range(2):
            assert R.relmax(y[b], want[b]) <= 1e-5, ("forward", N, env, b)
        wantb = ref.transform_batch(N, 1, want[:1], 1, True)
        gotb = s.transform_batch(torch.from_numpy(want[:1].copy()).cuda(), pf.PFFFT_BACKWARD, True).cpu().numpy()
        assert R.relmax(gotb[0], wantb[0]) <= 1e-5, ("backward", N, env)
        # round trip (tests/test_pffft.c:229-243): sum of squared errors <= N * 1e-7 per transform
        err = ((zd / N - xd) ** 2).sum(dim=1).max().item()
        assert err <= N * 1e-7, err
        # in place == out of place, bit for bit (bench_pffft.c:343-349)
        xi = xd.clone()
        s.transform_batch(xi, pf.PFFFT_FORWARD, True, out=xi)
        torch.cuda.synchronize()
        assert torch.equal(xi, yd)
    finally:
        s.close()


@pytest.mark.parametrize("N,env,name", [VARIANTS[0], VARIANTS[1], VARIANTS[4], VARIANTS[5]])
def test_cluster_persistent_loop_position_independent(pf, N, env, name):
    """more transforms than co-resident clusters: every cluster loops (barrier phases must stay matched), and the
    result of a transform must not depend on which cluster / iteration produced it (bit-exact)"""
    import torch
    s = make_setup(pf, N, 1, env, name)
    try:
        batch = max(8, (96 << 20) // (8 * N) + 3)           # ~96 MiB: several iterations for every cluster
        g = torch.Generator(device="cuda"); g.manual_seed(N)
        base = torch.rand((7, 2 * N), generator=g, device="cuda") * 2 - 1
        x = base.repeat((batch + 6) // 7, 1)[:batch].contiguous()
        y = s.transform_batch(x, pf.PFFFT_FORWARD, True)
        torch.cuda.synchronize()
        first = y[:7]
        for r in range(7, batch - 6, 7 * 13):
            assert torch.equal(y[r:r + 7], first), r
        tail = batch - (batch % 7) - 7
        assert torch.equal(y[tail:tail + 7], first)
        # canary: nothing written past the batch
        buf = torch.full((3 * 2 * N + 64,), 7.5, device="cuda")
        s.transform_batch(x[:3], pf.PFFFT_FORWARD, True, out=buf[:3 * 2 * N].view(3, 2 * N))
        torch.cuda.synchronize()
        assert bool((buf[3 * 2 * N:] == 7.5).all())
    finally:
        s.close()


@pytest.mark.parametrize("N", [32768, 131072])
def test_real_and_zdomain_wrap_the_cluster_kernel(pf, ref, R, N):
    """real N = 2 x Nc and the z-domain layouts run the cluster kernel between the generic load / store passes"""
    import torch
    s = make_setup(pf, N, 0, {"PFFFT_B200_CLUSTER": "all"}, "cluster")
    try:
        rng = np.random.default_rng(N)
        x = uniform(rng, 2 * N).reshape(2, N)
        xd = torch.from_numpy(x).cuda()
        y = s.transform_batch(xd, pf.PFFFT_FORWARD, True).cpu().numpy()
        want = ref.transform_batch(N, 0, x, 0, True)
        for b in range(2):
            assert R.relmax(y[b], want[b]) <= 1e-5
        yz = s.transform_batch(xd, pf.PFFFT_FORWARD, False)
        wz = ref.transform_batch(N, 0, x, 0, False)
        assert R.relmax(yz.cpu().numpy()[0], wz[0]) <= 1e-5
        back = s.transform_batch(yz, pf.PFFFT_BACKWARD, False).cpu().numpy()
        assert R.relmax(back[0] / N, x[0]) <= 1e-5
        assert np.array_equal(s.zreorder_batch(yz, pf.PFFFT_FORWARD).cpu().numpy(),
                              s.transform_batch(xd, pf.PFFFT_FORWARD, True).cpu().numpy())
    finally:
        s.close()


@pytest.mark.parametrize("N", [16384, 36864, 65536])
def test_cluster_switch_off_falls_back_to_two_pass(pf, ref, R, N):
    import torch
    with env_set({"PFFFT_B200_CLUSTER": "0"}):
        s = pf.Setup(N, 1)
        try:
            assert s.kernel.startswith("split_"), s.kernel
            x = uniform(np.random.default_rng(5), 3 * 2 * N).reshape(3, 2 * N)
            xd = torch.from_numpy(x).cuda()
            y = s.transform_batch(xd, 0, True)
            want = ref.transform_batch(N, 1, x[:1], 0, True)
            assert R.relmax(y[0].cpu().numpy(), want[0]) <= 1e-5
            z = s.transform_batch(y, 1, True)
            assert float(((z / N - xd) ** 2).sum(dim=1).max()) <= N * 1e-7
            xi = xd.clone()
            s.transform_batch(xi, 0, True, out=xi)
            assert torch.equal(xi, y)
        finally:
            s.close()
