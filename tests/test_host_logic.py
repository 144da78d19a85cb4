"""CPU-side checks of logic shared with the device code.  tests/emu/libemu.so (built by
__graft_entry__.build()) steps the PF_HD functions of pffft_b200/csrc -- index maps, Stockham stages,
real pre/post rotations, the warp kernel's two phases -- one lane at a time on the host, so the algebra
is verified against the reference before any GPU time is spent.  It is a test harness: the product
library never executes these on the CPU."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT, uniform

EMU = os.path.join(ROOT, "tests", "emu", "libemu.so")
L_C_ORD, L_C_Z, L_R_TIME, L_R_ORD, L_R_Z = range(5)
S_C_ORD, S_C_Z, S_R_TIME, S_R_ORD, S_R_Z = range(5)


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(EMU):
        pytest.skip("tests/emu/libemu.so not built (python __graft_entry__.py)")
    e = C.CDLL(EMU)
    e.emu_generic.argtypes = [C.c_int] * 6 + [C.c_void_p, C.c_void_p] + [C.c_longlong] * 4 + [C.c_int]
    e.emu_w1024.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_longlong]
    e.emu_fastconv_produced.restype = C.c_longlong
    e.emu_fastconv_produced.argtypes = [C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int]
    return e


def test_zdomain_index_maps_equal_reference_zreorder(emu, ref):
    for tr in (0, 1):
        for N in [32, 64, 96, 128, 160, 192, 256, 288, 384, 480, 512, 640, 800, 1024, 2592, 4000, 4096, 12000]:
            if not ref.lib.pffft_is_valid_size(N, tr):
                continue
            per = N if tr == 0 else 2 * N
            canon = ref.zreorder(N, tr, np.arange(per, dtype=np.float32), 0)   # canon[i] = z-domain index feeding slot i
            slots = N // 2 if tr == 0 else N
            pos = np.array([emu.emu_zpos(1 if tr == 0 else 0, k, N) for k in range(slots)])
            assert np.array_equal(canon[0::2], pos) and np.array_equal(canon[1::2], pos + 4), (N, tr)


def _emu_run(emu, prec, N, tr, d, lm, sm, x):
    dt = np.float32 if prec == 0 else np.float64
    n = N if tr == 0 else 2 * N
    x = np.ascontiguousarray(x, dtype=dt); o = np.zeros(n, dt)
    assert emu.emu_generic(prec, N, tr, d, lm, sm, x.ctypes.data, o.ctypes.data, 1, n, n, -1, n) == 0
    return o


@pytest.mark.parametrize("prec", [0, 1])
def test_generic_kernel_algebra(emu, ref, R, prec):
    dt = np.float32 if prec == 0 else np.float64
    rng = np.random.default_rng(1)
    for N in [16, 32, 48, 64, 96, 160, 240, 256, 480, 800, 1024, 2592, 4096]:
        for tr in (0, 1):
            if not ref.lib.pffft_is_valid_size(N, tr):
                continue
            tol = 2e-6 if prec == 0 else (1e-13 if (N & (N - 1)) == 0 else 2e-7)
            x = uniform(rng, N if tr == 0 else 2 * N, dt)
            want = ref.transform(N, tr, x, 0, True, dt); wantz = ref.transform(N, tr, x, 0, False, dt)
            fo = _emu_run(emu, prec, N, tr, 0, L_R_TIME if tr == 0 else L_C_ORD, S_R_ORD if tr == 0 else S_C_ORD, x)
            fz = _emu_run(emu, prec, N, tr, 0, L_R_TIME if tr == 0 else L_C_ORD, S_R_Z if tr == 0 else S_C_Z, x)
            bo = _emu_run(emu, prec, N, tr, 1, L_R_ORD if tr == 0 else L_C_ORD, S_R_TIME if tr == 0 else S_C_ORD, want)
            bz = _emu_run(emu, prec, N, tr, 1, L_R_Z if tr == 0 else L_C_Z, S_R_TIME if tr == 0 else S_C_ORD, wantz)
            assert R.relmax(fo, want) <= tol and R.relmax(fz, wantz) <= tol, (N, tr)
            assert R.relmax(bo, ref.transform(N, tr, want, 1, True, dt)) <= tol, (N, tr)
            assert R.relmax(bz, ref.transform(N, tr, wantz, 1, False, dt)) <= tol, (N, tr)


def test_register_fft_networks(emu):
    rng = np.random.default_rng(2)
    for N in (2, 4, 8, 16, 32, 64):
        for d in (0, 1):
            x = uniform(rng, 2 * N); o = np.zeros(2 * N, np.float32)
            assert emu.emu_regfft(N, d, x.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p)) == 0
            z = x[0::2].astype(np.float64) + 1j * x[1::2]
            w = np.fft.fft(z) if d == 0 else np.fft.ifft(z) * N
            assert np.max(np.abs((o[0::2] + 1j * o[1::2]) - w)) / np.max(np.abs(w)) <= 5e-7


def test_warp_kernel_phases_c1024(emu, ref, R):
    rng = np.random.default_rng(3)
    x = uniform(rng, 3 * 2048).reshape(3, 2048)
    for d in (0, 1):
        o = np.zeros_like(x)
        emu.emu_w1024(d, x.ctypes.data, o.ctypes.data, 3)
        w = ref.transform_batch(1024, 1, x, d, True)
        assert max(R.relmax(o[i], w[i]) for i in range(3)) <= 2e-6


def test_overlap_save_block_algebra_matches_reference_lengths(emu, ref):
    """the closed-form block plan used by pffastconv_apply yields exactly the reference loop's output count
    (tests/test_pffastconv.c:810-820 treats length mismatches as the hard failure)"""
    h = np.ones(200, np.float32)
    for taps in (1, 2, 31, 124, 131, 144, 200):
        for flags in (0, 1, 17):
            for bl in (0, 64, 512, 4096):
                for length in (taps - 1, taps, taps + 1, 300, 1000, 4097, 10000):
                    if length <= 0:
                        continue
                    cf = 2 if flags == 17 else 1
                    x = np.zeros(length * (2 if flags & 1 else 1) , np.float32)
                    for flush in (0, 1):
                        _, n_ref, bl_ref = ref.fastconv(h[:taps], x, bl, flags, flush)
                        nfft = bl_ref * cf
                        flen = 2 * taps - 1 if cf == 2 else taps
                        got = emu.emu_fastconv_produced(cf * length, nfft, flen, flush, 1 if cf == 2 else 0) // cf
                        assert got == n_ref, (taps, flags, bl, length, flush, got, n_ref)


def test_cta_kernel_phases_and_swizzle(emu, ref, R):
    """cta_kernels.cuh stepped on the CPU: all four sizes, complex and real (N/2 packing + pair rotation),
    canonical and z-domain; and the XOR-swizzled exchange tile audited bank-conflict free for every pass."""
    emu.emu_k2.argtypes = [C.c_int] * 5 + [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int]
    for c in (2, 4, 8, 16):
        assert emu.emu_k2_conflicts(c) == 1
    rng = np.random.default_rng(0)

    def run(Nc, N, lm, sm, d, x):
        x = np.ascontiguousarray(x); o = np.zeros_like(x)
        assert emu.emu_k2(Nc, N, lm, sm, d, x.ctypes.data, o.ctypes.data, -1, N) == 0
        return o
    for Nc in (512, 1024, 2048, 4096):
        N = Nc; x = uniform(rng, 2 * N)
        wf = ref.transform(N, 1, x, 0, True); wz = ref.transform(N, 1, x, 0, False)
        errs = [R.relmax(run(Nc, N, L_C_ORD, S_C_ORD, 0, x), wf), R.relmax(run(Nc, N, L_C_ORD, S_C_Z, 0, x), wz),
                R.relmax(run(Nc, N, L_C_ORD, S_C_ORD, 1, wf), ref.transform(N, 1, wf, 1, True)),
                R.relmax(run(Nc, N, L_C_Z, S_C_ORD, 1, wz), ref.transform(N, 1, wz, 1, False))]
        N = 2 * Nc; x = uniform(rng, N)
        wf = ref.transform(N, 0, x, 0, True); wz = ref.transform(N, 0, x, 0, False)
        errs += [R.relmax(run(Nc, N, L_R_TIME, S_R_ORD, 0, x), wf), R.relmax(run(Nc, N, L_R_TIME, S_R_Z, 0, x), wz),
                 R.relmax(run(Nc, N, L_R_ORD, S_R_TIME, 1, wf), ref.transform(N, 0, wf, 1, True)),
                 R.relmax(run(Nc, N, L_R_Z, S_R_TIME, 1, wz), ref.transform(N, 0, wz, 1, False))]
        assert max(errs) <= 2e-6, (Nc, errs)


def test_small_warp_kernel_phases(emu, ref, R):
    """N = 32..256 complex on the warp machinery (32/R2 transforms per warp chunk), including a partial last chunk"""
    emu.emu_wsmall.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_longlong]
    rng = np.random.default_rng(4)
    for N in (32, 64, 128, 256):
        for batch in (1, 1024 // N, 1024 // N + 3):
            x = uniform(rng, batch * 2 * N).reshape(batch, 2 * N)
            pad = np.zeros(((-(batch * N) % 1024) * 2,), np.float32)      # the emulation indexes whole 1024-point chunks
            xin = np.concatenate([x.ravel(), pad]); o = np.full_like(xin, np.nan)
            for d in (0, 1):
                assert emu.emu_wsmall(N, d, xin.ctypes.data, o.ctypes.data, batch) == 0
                w = ref.transform_batch(N, 1, x, d, True)
                got = o[: batch * 2 * N].reshape(batch, 2 * N)
                assert max(R.relmax(got[i], w[i]) for i in range(batch)) <= 2e-6, (N, batch, d)
                assert np.all(np.isnan(o[batch * 2 * N:])), "wrote beyond the batch"


def test_mixed_radix_warp_kernel_phases(emu, ref, R):
    """N = 32*R2, R2 in {3,5,6,9,10,12,15}: register radix-3/5 DFTs + radix-32 columns, stepped on the CPU"""
    emu.emu_wmixed.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_longlong]
    rng = np.random.default_rng(5)
    for r2 in (3, 5, 6, 9, 10, 12, 15, 18, 20, 24, 25, 27, 30):
        N = 32 * r2
        tw = 32 // r2
        for batch in (1, tw, tw + 2):
            x = uniform(rng, batch * 2 * N).reshape(batch, 2 * N)
            nchunks = -(-batch // tw)
            xin = np.zeros(nchunks * tw * 2 * N, np.float32); xin[: x.size] = x.ravel()
            o = np.full_like(xin, np.nan)
            for d in (0, 1):
                assert emu.emu_wmixed(N, d, xin.ctypes.data, o.ctypes.data, batch) == 0
                w = ref.transform_batch(N, 1, x, d, True)
                got = o[: batch * 2 * N].reshape(batch, 2 * N)
                assert max(R.relmax(got[i], w[i]) for i in range(batch)) <= 2e-6, (N, batch, d)
                assert np.all(np.isnan(o[batch * 2 * N:])), "wrote beyond the batch"


def test_cluster_kernel_phases(emu, ref, R):
    """cluster plans (cluster_kernels.cuh): the CTAs of one cluster stepped phase by phase on the CPU, cluster barriers =
    phase boundaries, DSMEM = the peers' buffers.  Every shape / row mode that is instantiated, against the reference."""
    emu.emu_cluster.argtypes = [C.c_int] * 4 + [C.c_void_p] * 2
    rng = np.random.default_rng(11)
    for CL, Q, scatter in [(2, 1, 0), (4, 1, 0), (4, 1, 1), (8, 1, 1), (8, 2, 0), (4, 4, 0), (16, 1, 0), (16, 1, 1)]:
        N = CL * Q * 4096
        x = (rng.random(2 * N) * 2 - 1).astype(np.float32)
        want = ref.transform(N, 1, x, 0, True)
        for d in ((0, 1) if CL == 4 else (0,)):
            o = np.zeros(2 * N, np.float32)
            src = x if d == 0 else want
            assert emu.emu_cluster(CL, Q, scatter, d, src.ctypes.data, o.ctypes.data) == 0
            if d == 0:
                assert R.relmax(o, want) <= 2e-6, (CL, Q, scatter)
            else:
                assert R.relmax(o / N, x) <= 2e-6, (CL, Q, scatter)


def test_single_cta_two_level_kernel_phases(emu, ref, R):
    """k_cta_split with the row-major combine twiddles: (C, R) = rows of 256*C points x radix-R finish"""
    emu.emu_cta_split.argtypes = [C.c_int] * 3 + [C.c_void_p] * 2
    rng = np.random.default_rng(12)
    for c, r in [(16, 2), (8, 3), (4, 9), (2, 15), (8, 6)]:
        N = r * 256 * c
        x = (rng.random(2 * N) * 2 - 1).astype(np.float32)
        o = np.zeros(2 * N, np.float32)
        assert emu.emu_cta_split(c, r, 0, x.ctypes.data, o.ctypes.data) == 0
        assert R.relmax(o, ref.transform(N, 1, x, 0, True)) <= 2e-6, (c, r)


def test_tiled_2d_large_n_phases(emu, ref, R):
    """tiled two-dimensional plan (tiled2d_kernels.cuh, opt-in): pass A tiles then pass C tiles stepped thread by thread;
    both exchange tiles must be bank-conflict free for every shape"""
    emu.emu_t2d.argtypes = [C.c_int] * 3 + [C.c_void_p] * 2
    rng = np.random.default_rng(13)
    for a1, a2 in [(8, 8), (16, 8), (8, 16), (16, 16)]:
        assert emu.emu_t2d_conflicts(a1, a2) == 1
        N = 256 * a1 * a2
        x = (rng.random(2 * N) * 2 - 1).astype(np.float32)
        want = ref.transform(N, 1, x, 0, True)
        o = np.zeros(2 * N, np.float32)
        assert emu.emu_t2d(a1, a2, 0, x.ctypes.data, o.ctypes.data) == 0
        assert R.relmax(o, want) <= 2e-6, (a1, a2)
        assert emu.emu_t2d(a1, a2, 1, want.ctypes.data, o.ctypes.data) == 0
        assert R.relmax(o / N, x) <= 2e-6, (a1, a2)


def test_tiled_2d_cluster_fused_phases(emu, ref, R):
    """cluster-fused form of the tiled plan (pass A -> pass C through the peers' shared memory): every cluster shape"""
    emu.emu_t2d_cluster.argtypes = [C.c_int] * 4 + [C.c_void_p] * 2
    rng = np.random.default_rng(14)
    for a1, a2, cl in [(8, 8, 8), (8, 8, 4), (16, 8, 8), (16, 16, 8), (16, 16, 16)]:
        N = 256 * a1 * a2
        x = (rng.random(2 * N) * 2 - 1).astype(np.float32)
        want = ref.transform(N, 1, x, 0, True)
        o = np.zeros(2 * N, np.float32)
        assert emu.emu_t2d_cluster(a1, a2, cl, 0, x.ctypes.data, o.ctypes.data) == 0
        assert R.relmax(o, want) <= 2e-6, (a1, a2, cl)
        if cl == 8:
            assert emu.emu_t2d_cluster(a1, a2, cl, 1, want.ctypes.data, o.ctypes.data) == 0
            assert R.relmax(o / N, x) <= 2e-6, (a1, a2, cl)


def test_tiled_2d_general_radix_phases(emu, ref, R):
    """general-radix tiled plan (Nc = 256*A1*A2, radix 3/5 factors included): every instantiated shape, conflict audit"""
    emu.emu_t2dg.argtypes = [C.c_int] * 3 + [C.c_void_p] * 2
    rng = np.random.default_rng(15)
    for a1, a2 in [(6, 5), (6, 6), (8, 6), (10, 8), (12, 8), (12, 12), (16, 10), (16, 12), (16, 15), (8, 8), (16, 16)]:
        assert emu.emu_t2dg_conflicts(a1, a2) == 1, (a1, a2)
        N = 256 * a1 * a2
        x = (rng.random(2 * N) * 2 - 1).astype(np.float32)
        want = ref.transform(N, 1, x, 0, True)
        o = np.zeros(2 * N, np.float32)
        assert emu.emu_t2dg(a1, a2, 0, x.ctypes.data, o.ctypes.data) == 0
        assert R.relmax(o, want) <= 2e-6, (a1, a2)
        if a1 in (6, 12):
            assert emu.emu_t2dg(a1, a2, 1, want.ctypes.data, o.ctypes.data) == 0
            assert R.relmax(o / N, x) <= 2e-6, (a1, a2)


def test_tiled_2d_double_phases(emu):
    """the double-precision instantiations of the general tiled plan against numpy float64 (1e-13; north_star: 1e-12)"""
    emu.emu_t2dg_double.argtypes = [C.c_int] * 3 + [C.c_void_p] * 2
    rng = np.random.default_rng(16)
    for a1, a2 in [(8, 8), (16, 8), (16, 16)]:
        N = 256 * a1 * a2
        x = rng.random(2 * N) * 2 - 1
        o = np.zeros(2 * N)
        assert emu.emu_t2dg_double(a1, a2, 0, x.ctypes.data, o.ctypes.data) == 0
        want = np.fft.fft(x.view(np.complex128))
        assert np.abs(o.view(np.complex128) - want).max() <= 1e-13 * np.abs(want).max(), (a1, a2)


def test_tiled_2d_carrier_dynamic_range(emu):
    """the reference's pure-carrier property (tests/test_pffft.c:109-213: spur-free range >= 140 dB, phase within 1e-4
    degree, magnitude within 1e-6) on the CPU-stepped arithmetic of the tiled plans that are the default for 32768 / 65536"""
    emu.emu_t2d.argtypes = [C.c_int] * 3 + [C.c_void_p] * 2
    for a1, a2 in [(16, 8), (16, 16)]:
        N = 256 * a1 * a2
        for m, k in enumerate(range(0, N, N // 8)):
            amp = 1.0 if m % 3 == 0 else 1.1
            freq = k / N if k < N / 2 else (k - N) / N
            dphi = 2 * np.pi * freq
            if dphi < 0:
                dphi += 2 * np.pi
            phi0 = (m % 4) * 0.125 * np.pi
            # the reference's normalised phase accumulation, vectorised: phi_j = wrap(phi0 + j*dphi) computed incrementally
            ph = np.empty(N)
            phi = phi0
            for j in range(N):
                ph[j] = phi
                phi += dphi
                if phi >= np.pi:
                    phi -= 2 * np.pi
            x = np.stack([amp * np.cos(ph).astype(np.float32), amp * np.sin(ph).astype(np.float32)], -1).ravel().astype(np.float32)
            o = np.zeros_like(x)
            assert emu.emu_t2d(a1, a2, 0, x.ctypes.data, o.ctypes.data) == 0
            y = o.astype(np.float64)
            p = y[0::2] ** 2 + y[1::2] ** 2
            dyn = 10 * np.log10(p[k]) - 10 * np.log10(max(np.delete(p, k).max(), 1e-300))
            assert dyn >= 140.0, (N, k, dyn)
            assert abs(np.sqrt(p[k]) / N - amp) <= 1e-6, (N, k)
            if k > 0 and k != N // 2:
                assert abs(np.arctan2(y[2 * k + 1], y[2 * k]) - phi0) <= 1e-4 * np.pi / 180, (N, k)
