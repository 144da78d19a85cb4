// cta_hooks.cuh -- plan tables and launchers of the 16x16xC CTA kernels, shared by the float and double APIs.
#pragma once
#include "engine.cuh"
#include "cta_kernels.cuh"

namespace pf {

inline int cta_C_for(int Nc) { return Nc == 512 ? 2 : Nc == 1024 ? 4 : Nc == 2048 ? 8 : Nc == 4096 ? 16 : 0; }
inline const char* cta_name(int C) { return C == 2 ? "cta_16x16x2" : C == 4 ? "cta_16x16x4" : C == 8 ? "cta_16x16x8" : "cta_16x16x16"; }
inline size_t cta_table_cpx(int Nc) { const int C = cta_C_for(Nc); return C ? (size_t)Nc + 16 * (size_t)C : 0; }

// [tw1: Nc][tw2: 16*C]   tw1[ka*BC + m] = exp(-2 pi i m ka / Nc),  tw2[kb*C + nc] = exp(-2 pi i nc kb / BC)
template <typename T> void cta_fill_tables(int Nc, T* dst) {
  const int C = cta_C_for(Nc);
  if (!C) return;
  const int BC = 16 * C;
  for (int ka = 0; ka < 16; ++ka)
    for (int m = 0; m < BC; ++m) {
      long double c, sn;
      pfplan::unit_root((long long)m * ka, Nc, &c, &sn);
      dst[2 * (ka * BC + m)] = (T)c; dst[2 * (ka * BC + m) + 1] = (T)sn;
    }
  T* t2 = dst + 2 * (size_t)Nc;
  for (int kb = 0; kb < 16; ++kb)
    for (int nc = 0; nc < C; ++nc) {
      long double c, sn;
      pfplan::unit_root((long long)nc * kb, BC, &c, &sn);
      t2[2 * (kb * C + nc)] = (T)c; t2[2 * (kb * C + nc) + 1] = (T)sn;
    }
}

// threads per SM the register budget is sized for: 1024 (64 regs) float, 512 (128 regs) double
template <typename T> constexpr int cta_tpsm() { return sizeof(T) == 4 ? 1024 : 512; }

template <typename T, int C, int LM, int SM, int SIGN, bool STAGED>
int launch_cta_v(Setup<T>* s, const XformParams<T>& p, cudaStream_t st) {
  constexpr int MINB = cta_tpsm<T>() / (16 * C);
  auto kern = k_cta_fft<T, C, LM, SM, SIGN, MINB, STAGED>;
  // second buffer: TMA stage, gather staging of z-domain / backward-real inputs, or the z image of a C=16 real forward
  constexpr bool kSecond = STAGED || LM == L_C_Z || LM == L_R_Z || (SM == S_R_Z && C == 16);
  const size_t smem = (size_t)K2<C>::NC * sizeof(cpx<T>) * (kSecond ? 2 : 1) + (STAGED ? 16 : 0);
  static thread_local int per_sm = 0;
  if (per_sm == 0) {
    if (smem > 48 * 1024) PF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 16 * C, smem);
    if (per_sm < 1) per_sm = 1;
  }
  long long ctas = p.batch;
  const long long cap = (long long)s->sm_count * per_sm;
  if (ctas > cap) ctas = cap;
  const cpx<T>* tw1 = s->tw_fast;
  const cpx<T>* tw2 = s->tw_fast + K2<C>::NC;
  kern<<<(int)ctas, 16 * C, smem, st>>>(p, tw1, tw2);
  count_launch();
  PF_CUDA_OK(cudaGetLastError());
  return 0;
}
// TMA-staged inputs measured slower than register-fed loads at 1024 threads/SM (C3: 0.70 vs 0.79 of HBM peak):
// opt-in with PFFFT_B200_CTA_STAGE=1, float only, contiguous 16-byte aligned canonical input
inline bool cta_stage_enabled() {
  static const bool on = getenv("PFFFT_B200_CTA_STAGE") ? atoi(getenv("PFFFT_B200_CTA_STAGE")) != 0 : false;
  return on;
}
template <typename T, int C, int LM, int SM, int SIGN>
int launch_cta(Setup<T>* s, const XformParams<T>& p, cudaStream_t st) {
  if constexpr (sizeof(T) == 4 && (LM == L_C_ORD || LM == L_R_TIME)) {
    const bool contiguous = p.in_limit < 0 && p.in_stride == (long long)s->per() && (reinterpret_cast<uintptr_t>(p.in) & 15) == 0;
    if (cta_stage_enabled() && contiguous) return launch_cta_v<T, C, LM, SM, SIGN, true>(s, p, st);
  }
  return launch_cta_v<T, C, LM, SM, SIGN, false>(s, p, st);
}
template <typename T, int C>
int run_cta(Setup<T>* s, const XformParams<T>& p, int direction, int ordered, cudaStream_t st) {
  const bool fwd = direction == DIR_FORWARD;
  if (s->transform == XF_COMPLEX) {
    if (fwd) return ordered ? launch_cta<T, C, L_C_ORD, S_C_ORD, -1>(s, p, st) : launch_cta<T, C, L_C_ORD, S_C_Z, -1>(s, p, st);
    return ordered ? launch_cta<T, C, L_C_ORD, S_C_ORD, +1>(s, p, st) : launch_cta<T, C, L_C_Z, S_C_ORD, +1>(s, p, st);
  }
  if (fwd) return ordered ? launch_cta<T, C, L_R_TIME, S_R_ORD, -1>(s, p, st) : launch_cta<T, C, L_R_TIME, S_R_Z, -1>(s, p, st);
  return ordered ? launch_cta<T, C, L_R_ORD, S_R_TIME, +1>(s, p, st) : launch_cta<T, C, L_R_Z, S_R_TIME, +1>(s, p, st);
}
template <typename T>
int run_cta_any(Setup<T>* s, int C, const XformParams<T>& p, int direction, int ordered, cudaStream_t st) {
  switch (C) {
    case 2: return run_cta<T, 2>(s, p, direction, ordered, st);
    case 4: return run_cta<T, 4>(s, p, direction, ordered, st);
    case 8: return run_cta<T, 8>(s, p, direction, ordered, st);
    default: return run_cta<T, 16>(s, p, direction, ordered, st);
  }
}


// ---------------------------------------------------------------- large N: Nc = R x 4096, R in {2,4,8,16}
// (complex N = 8192..65536, real N = 16384..131072).  Decimation in time over the first digit:
//   rows   : R launches of the 16x16x16 CTA kernel, row n1 = FFT_4096 of x[n1 + R*n2] (element stride R)
//   combine: k_split_combine finishes with the radix-R butterflies and writes X in natural order.
// Two HBM round trips (plus one for a real pre-/post-rotation or z-domain pass): ceiling 0.5 of the roofline,
// against (stages+2) round trips of the global Stockham path it replaces.
inline int split_R_for(int Nc) { return Nc == 8192 ? 2 : Nc == 16384 ? 4 : Nc == 32768 ? 8 : Nc == 65536 ? 16 : 0; }
inline const char* split_name(int R) { return R == 2 ? "split_2x4096" : R == 4 ? "split_4x4096" : R == 8 ? "split_8x4096" : "split_16x4096"; }

template <typename T, int SIGN>
int split_core(Setup<T>* s, int R, const cpx<T>* src, cpx<T>* rows, cpx<T>* dst, long long batch, cudaStream_t st) {
  constexpr int N2 = 4096;
  // ONE launch over batch*R rows: consecutive CTAs take the R interleaved sub-sequences of the same transform, so
  // their stride-R reads of the same 128-byte lines meet in L2 instead of re-reading DRAM R times
  XformParams<T> q;
  q.in = reinterpret_cast<const T*>(src); q.out = reinterpret_cast<T*>(rows);
  q.in_stride = 2LL * s->Nc; q.in_group = R; q.in_gstep = 2; q.in_estride = R;
  q.out_stride = 2LL * N2; q.in_limit = -1; q.out_count = 2 * N2;
  q.batch = batch * R; q.N = N2; q.Nc = N2; q.nfac = 0; q.tw = s->tw; q.twr = nullptr;
  for (int i = 0; i < PF_MAX_FACTORS; ++i) { q.fac[i] = 1; q.magic[i] = 0; }
  q.magic_nc = 0;
  { const int rc = launch_cta_v<T, 16, L_C_ORD, S_C_ORD, SIGN, false>(s, q, st); if (rc) return rc; }
  const long long work = batch * N2;
  long long g = (work + 255) / 256; const long long cap = (long long)s->sm_count * 16;
  if (g > cap) g = cap; if (g < 1) g = 1;
  switch (R) {
    case 2: k_split_combine<T, 2, SIGN><<<(int)g, 256, 0, st>>>(rows, dst, batch, N2, s->tw); break;
    case 4: k_split_combine<T, 4, SIGN><<<(int)g, 256, 0, st>>>(rows, dst, batch, N2, s->tw); break;
    case 8: k_split_combine<T, 8, SIGN><<<(int)g, 256, 0, st>>>(rows, dst, batch, N2, s->tw); break;
    default: k_split_combine<T, 16, SIGN><<<(int)g, 256, 0, st>>>(rows, dst, batch, N2, s->tw); break;
  }
  count_launch();
  PF_CUDA_OK(cudaGetLastError());
  return 0;
}

template <typename T, int LM, int SM, int SIGN>
int run_split_modes(Setup<T>* s, int R, const XformParams<T>& p, cudaStream_t st) {
  std::lock_guard<std::mutex> lock(s->scratch_mu);
  { const int rc = scratch_acquire(s, (size_t)p.batch * s->Nc, st); if (rc) return rc; }
  const long long total = p.batch * (long long)s->Nc;
  long long g = (total + 255) / 256; const long long cap = (long long)s->sm_count * 32;
  if (g > cap) g = cap; if (g < 1) g = 1;
  const bool dense_io = p.in_stride == (long long)s->per() && p.out_stride == (long long)s->per();
  // input that already IS the dense complex core (complex canonical, or real time samples read as pairs)
  const bool direct_in = dense_io && p.in_limit < 0 && (LM == L_C_ORD || LM == L_R_TIME) && vec_aligned<T>(p.in);
  const bool direct_out = dense_io && (SM == S_C_ORD || (SM == S_R_TIME && p.out_count >= s->N)) && vec_aligned<T>(p.out);
  const cpx<T>* src = reinterpret_cast<const cpx<T>*>(p.in);
  if (!direct_in) {
    k_glob_load<T, LM><<<(int)g, 256, 0, st>>>(p, s->d_scratch[0]);
    count_launch();
    src = s->d_scratch[0];
  }
  cpx<T>* dst = direct_out ? reinterpret_cast<cpx<T>*>(p.out) : s->d_scratch[0];
  const int rc = split_core<T, SIGN>(s, R, src, s->d_scratch[1], dst, p.batch, st);
  if (rc) return rc;
  if (!direct_out) {
    k_glob_store<T, SM><<<(int)g, 256, 0, st>>>(p, s->d_scratch[0]);
    count_launch();
  }
  PF_CUDA_OK(cudaGetLastError());
  PF_CUDA_OK(cudaEventRecord(s->scratch_done, st));
  return 0;
}
template <typename T>
int run_split(Setup<T>* s, int R, const XformParams<T>& p, int direction, int ordered, cudaStream_t st) {
  const bool fwd = direction == DIR_FORWARD;
  if (s->transform == XF_COMPLEX) {
    if (fwd) return ordered ? run_split_modes<T, L_C_ORD, S_C_ORD, -1>(s, R, p, st) : run_split_modes<T, L_C_ORD, S_C_Z, -1>(s, R, p, st);
    return ordered ? run_split_modes<T, L_C_ORD, S_C_ORD, +1>(s, R, p, st) : run_split_modes<T, L_C_Z, S_C_ORD, +1>(s, R, p, st);
  }
  if (fwd) return ordered ? run_split_modes<T, L_R_TIME, S_R_ORD, -1>(s, R, p, st) : run_split_modes<T, L_R_TIME, S_R_Z, -1>(s, R, p, st);
  return ordered ? run_split_modes<T, L_R_ORD, S_R_TIME, +1>(s, R, p, st) : run_split_modes<T, L_R_Z, S_R_TIME, +1>(s, R, p, st);
}

// hooks for a precision whose only tuned kernels are the CTA ones (double)
template <typename T> struct CtaOnlyHooks {
  static size_t extra_table_cpx(int N, int transform) {
    const int Nc = transform == XF_REAL ? N / 2 : N;
    return split_R_for(Nc) ? cta_table_cpx(4096) : cta_table_cpx(Nc);
  }
  static void fill_extra_table(int N, int transform, T* dst) {
    const int Nc = transform == XF_REAL ? N / 2 : N;
    cta_fill_tables<T>(split_R_for(Nc) ? 4096 : Nc, dst);
  }
  static bool plan(Setup<T>* s) {
    if (const int R = split_R_for(s->Nc)) {
      if (getenv("PFFFT_B200_NO_SPLIT")) return false;
      s->fast_variant = 300 + R;
      s->kernel_name = split_name(R);
      return true;
    }
    const int C = cta_C_for(s->Nc);
    if (!C || getenv("PFFFT_B200_NO_CTA")) return false;
    s->fast_variant = 100 + C;
    s->kernel_name = cta_name(C);
    return true;
  }
  static int run(Setup<T>* s, const T* in, T* out, long long batch, int direction, int ordered, cudaStream_t st, const XformOpts& o) {
    const XformParams<T> p = make_params(s, in, out, batch, o);
    if (s->fast_variant >= 300) return run_split<T>(s, s->fast_variant - 300, p, direction, ordered, st);
    return run_cta_any<T>(s, s->fast_variant - 100, p, direction, ordered, st);
  }
};

}  // namespace pf
