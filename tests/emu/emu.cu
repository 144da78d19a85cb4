// emu.cu -- DEVELOPMENT HARNESS, not part of the product.
// Steps the PF_HD device functions of pffft_b200/csrc lane-by-lane on the CPU so index math and
// butterfly algebra can be checked in the GPU-less build container before a gpurun call.
// Nothing in pffft_b200/ links or loads this; only tests/test_emu_*.py does.
#include <vector>
#include <cstring>
#include "../../pffft_b200/csrc/generic_kernels.cuh"
#include "../../pffft_b200/csrc/plan.h"
#include "emu_fast.cuh"

using namespace pf;

template <typename T, int LM, int SM, int SIGN>
static void run_generic(XformParams<T> p) {
  std::vector<cpx<T>> A(p.Nc), B(p.Nc);
  for (long long t = 0; t < p.batch; ++t) {
    const T* ibase = p.in + t * p.in_stride;
    const long long avail = p.in_limit < 0 ? -1 : p.in_limit - t * p.in_stride;
    for (int i = 0; i < p.Nc; ++i) A[i] = load_core<LM, T>(ibase, i, p.N, p.Nc, p.twr, avail, vec_aligned<T>(ibase));
    cpx<T>* src = A.data(); cpx<T>* dst = B.data();
    int s = 1;
    for (int f = 0; f < p.nfac; ++f) {
      const int r = p.fac[f], m = p.Nc / r;
      for (int li = 0; li < 4; ++li) stockham_stage<SIGN, T>(r, src, dst, li, 4, p.Nc, s, p.magic[f], p.tw);   // 4 emulated lanes
      std::swap(src, dst); s *= r;
    }
    T* obase = p.out + t * p.out_stride;
    std::vector<cpx<T>> Z(src, src + p.Nc);   // store may alias input
    for (int k = 0; k < p.Nc; ++k) store_core<SM, T>(obase, Z.data(), k, p.N, p.Nc, p.twr, p.out_count, vec_aligned<T>(obase));
  }
}

template <typename T, int LM, int SM>
static void run_dir(XformParams<T> p, int dir) { if (dir == 0) run_generic<T, LM, SM, -1>(p); else run_generic<T, LM, SM, +1>(p); }

template <typename T>
static int emu_generic_t(int N, int transform, int dir, int lm, int sm, const T* in, T* out, long long batch,
                         long long in_stride, long long out_stride, long long in_limit, int out_count) {
  if (!pfplan::setup_size_ok(N, transform)) return -1;
  XformParams<T> p{};
  p.N = N; p.Nc = transform == 0 ? N / 2 : N;
  auto f = pfplan::factorize(p.Nc);
  p.nfac = (int)f.size(); { int prod = 1; for (int i = 0; i < p.nfac; ++i) { p.fac[i] = f[i]; p.magic[i] = stage_magic(prod); prod *= f[i]; } }
  std::vector<T> tw(2 * (size_t)p.Nc), twr(2 * (size_t)(N / 2));
  pfplan::fill_roots<T>(tw.data(), p.Nc, p.Nc);
  pfplan::fill_roots<T>(twr.data(), N / 2, N);
  p.tw = reinterpret_cast<const cpx<T>*>(tw.data()); p.twr = reinterpret_cast<const cpx<T>*>(twr.data());
  p.in = in; p.out = out; p.batch = batch; p.in_stride = in_stride; p.out_stride = out_stride;
  p.in_limit = in_limit; p.out_count = out_count;
#define CASE(L, S) if (lm == L && sm == S) { run_dir<T, L, S>(p, dir); return 0; }
  CASE(L_C_ORD, S_C_ORD) CASE(L_C_ORD, S_C_Z) CASE(L_C_Z, S_C_ORD) CASE(L_C_Z, S_C_Z)
  CASE(L_R_TIME, S_R_ORD) CASE(L_R_TIME, S_R_Z) CASE(L_R_ORD, S_R_TIME) CASE(L_R_Z, S_R_TIME)
#undef CASE
  return -2;
}

extern "C" int emu_generic(int prec, int N, int transform, int dir, int lm, int sm, const void* in, void* out,
                           long long batch, long long in_stride, long long out_stride, long long in_limit, int out_count) {
  if (prec == 0) return emu_generic_t<float>(N, transform, dir, lm, sm, (const float*)in, (float*)out, batch, in_stride, out_stride, in_limit, out_count);
  return emu_generic_t<double>(N, transform, dir, lm, sm, (const double*)in, (double*)out, batch, in_stride, out_stride, in_limit, out_count);
}

// z-domain index maps, for a direct comparison with the reference's pffft_zreorder
extern "C" int emu_zpos(int real, int k, int N) { return real ? zpos_real(k, N) : zpos_complex(k, N); }

// register FFT check: x[N] natural order in, X[N] natural order out
template <int N, int SIGN> static void regfft_run(const float* in, float* out) {
  cpx<float> v[N];
  constexpr int bits = ct::ilog2(N);
  for (int p = 0; p < N; ++p) { int n = ct::bitrev(p, bits); v[p] = mk<float>(in[2 * n], in[2 * n + 1]); }
  reg_fft<N, SIGN>(v);
  for (int k = 0; k < N; ++k) { out[2 * k] = v[k].x; out[2 * k + 1] = v[k].y; }
}
extern "C" int emu_regfft(int N, int dir, const float* in, float* out) {
#define RF(n) if (N == n) { if (dir == 0) regfft_run<n, -1>(in, out); else regfft_run<n, +1>(in, out); return 0; }
  RF(2) RF(4) RF(8) RF(16) RF(32) RF(64)
#undef RF
  return -1;
}

// overlap-save block algebra (host logic of pffastconv_apply) for a CPU check against the reference
extern "C" long long emu_fastconv_produced(long long inputLen, int Nfft, int filterLen, int flush, int even_out) {
  return pfplan::plan_blocks(inputLen, Nfft, filterLen, flush, even_out != 0).produced;
}
