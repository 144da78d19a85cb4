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
  const size_t smem = (size_t)K2<C>::NC * sizeof(cpx<T>) * (STAGED ? 2 : 1) + (STAGED ? 16 : 0);
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

// hooks for a precision whose only tuned kernels are the CTA ones (double)
template <typename T> struct CtaOnlyHooks {
  static size_t extra_table_cpx(int N, int transform) { return cta_table_cpx(transform == XF_REAL ? N / 2 : N); }
  static void fill_extra_table(int N, int transform, T* dst) { cta_fill_tables<T>(transform == XF_REAL ? N / 2 : N, dst); }
  static bool plan(Setup<T>* s) {
    const int C = cta_C_for(s->Nc);
    if (!C || getenv("PFFFT_B200_NO_CTA")) return false;
    s->fast_variant = 100 + C;
    s->kernel_name = cta_name(C);
    return true;
  }
  static int run(Setup<T>* s, const T* in, T* out, long long batch, int direction, int ordered, cudaStream_t st, const XformOpts& o) {
    const XformParams<T> p = make_params(s, in, out, batch, o);
    return run_cta_any<T>(s, s->fast_variant - 100, p, direction, ordered, st);
  }
};

}  // namespace pf
